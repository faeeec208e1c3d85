"""BASELINE.json configs 3, 4, 5 through the host drop-in on one GPU: build, render, time, sanity (GPU box)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host
from PIL import Image

which = sys.argv[1:] or ["c3", "c4", "c5"]
os.makedirs("gpurun_out", exist_ok=True)
for cfg in which:
    t0 = time.time()
    if cfg == "c3":
        sc = scenes.triangle_mesh(1_000_000); W, H, nit, cam, frames = 1920, 1080, 5, wire.CT_PERSPECTIVE, [0]
    elif cfg == "c4":
        sc = scenes.config4(); W, H, nit, cam, frames = 3840, 2160, 3, wire.CT_PERSPECTIVE, [10, 11, 12, 13]
    else:
        sc = scenes.config2(); W, H, nit, cam, frames = 3840, 2160, 3, wire.CT_ANAGLYPH, list(range(16))
    si = wire.default_scene_info(W, H, nb_ray_iterations=nit)
    si.cameraType = cam
    si.maxPathTracingIterations = 1 << 30
    h = host.SceneHost(si, limits=(max(W, 1920), max(H, 1080)), capacity=(16_000_000, 4_000_000))
    nb = sc.replay(h)
    t_build = time.time() - t0
    h.set_randoms(np.zeros(max(W, 1920) * max(H, 1080), np.float32), 0)
    h.set_camera(sc.eye, sc.target, sc.angles)
    h.init_buffers()
    lib = engine.load()
    if os.environ.get('SOLR_OPT6') is not None: lib.b200_set_option(6, int(os.environ['SOLR_OPT6']))
    if os.environ.get('SOLR_OPT4') is not None: lib.b200_set_option(4, int(os.environ['SOLR_OPT4']))
    e = engine.Engine.__new__(engine.Engine); e.lib = lib
    ms = []
    for it in frames:
        si.pathTracingIteration = it
        h.set_scene_info(si); h.set_camera(sc.eye, sc.target, sc.angles)
        t1 = time.time(); h.render_begin(0.0); h.render_end(); ms.append((time.time() - t1) * 1e3)
    e.check = lambda: None
    rays, px = e.counters(reset=True)
    bm = h.bitmap().copy(); ids = h.primitive_ids().copy()
    print("%s: prims %d boxes %d build %.1fs  stats %s" % (cfg, sc.nb_primitives, nb, t_build, e.scene_stats()))
    print("   frames %s: e2e ms/frame %s  kernel ms(last) %.2f  rays %d px %d  hit px %.1f%%  mean rgb %.1f" % (
        frames, ["%.1f" % m for m in ms], e.last_render_ms(), rays, px, 100.0 * (ids[..., 0] >= 0).mean(), bm.mean()))
    Image.fromarray(bm[::-1]).resize((960, 540)).save("gpurun_out/%s.png" % cfg)
    h.close()
