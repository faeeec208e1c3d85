"""Where an end-to-end frame's time goes (config 2 through the host drop-in, GPU box): render_begin (parameter upload + launches),
render_end (wait + read-back) and the kernels' own time (events), with the outputs copied after the kernels / streamed by them,
ids every frame / on demand."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, engine, host, workloads
key = sys.argv[1] if len(sys.argv) > 1 else "config2"
wl = workloads.WORKLOADS[key]
W, H = wl["size"]
sc = wl["scene"]()
si = workloads.scene_info(key)
h = host.SceneHost(si, limits=wl["limits"], capacity=wl["capacity"])
sc.replay(h)
h.set_randoms(np.zeros(max(wire.REF_MAX_BITMAP_SIZE, W * H), np.float32), 0)
h.set_camera(sc.eye, sc.target, sc.angles)
h.init_buffers()
lib = engine.load()
sl = h.scene_info
sl.maxPathTracingIterations = 1 << 30
its = wl["iterations"]
def frame(k):
    sl.pathTracingIteration = its[k % len(its)]
    h.set_scene_info(sl)
    t0 = time.perf_counter(); h.render_begin(0.0); t1 = time.perf_counter(); h.render_end(); t2 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, float(lib.b200_last_render_ms())
for lazy in (False, True):
    h.set_lazy_ids(lazy)
    for opt in (0, 1):
        lib.b200_set_option(12, opt)
        for k in range(4 * len(its)): frame(k)
        n0 = int(lib.b200_frames_streamed())
        r = np.array([frame(k) for k in range(10 * len(its))])
        print("%s ids %-9s outputs %-8s render_begin %.3f ms  render_end %.3f ms  sum %.3f ms  kernels (events) %.3f ms  frames streamed %d/%d" %
              (key, "on demand" if lazy else "eager", "streamed" if opt else "copied", r[:, 0].mean(), r[:, 1].mean(), r[:, :2].sum(1).mean(), r[:, 2].mean(),
               int(lib.b200_frames_streamed()) - n0, len(r)), flush=True)
h.close()
