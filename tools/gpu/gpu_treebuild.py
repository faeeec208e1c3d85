"""GPU tree build (option 10) against the host SAH build on config 2 (and config 4 with 'big'): upload milliseconds, nodes, frame ms, checksums."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host
for name in (["config2"] + (["config4", "config3"] if "big" in sys.argv else [])):
    sc = scenes.config2() if name == "config2" else scenes.config4() if name == "config4" else scenes.triangle_mesh()
    W, H = (1920, 1080)
    si = wire.default_scene_info(W, H, nb_ray_iterations=3)
    h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
    for gpu in (0, 1, 1):
        e = engine.Engine(si); e.set_option(10, gpu)
        t = time.time(); e.upload(a, randoms=np.zeros(W * H, np.float32)); up = time.time() - t
        st = e.scene_stats()
        ms = []
        for it in range(5):
            e.render(si, sc.eye, sc.target, sc.angles); e.synchronize(); ms.append(e.last_render_ms())
        bm, ids = e.readback(si)
        print("%s gpu=%d upload(all arrays) %.1f ms  h2d_scene %.1f ms  nodes %d + %d  frame %.3f ms  checksum %d %d" % (
            name, gpu, 1e3 * up, st["upload_ms"], st["walk_tree_nodes"], st["point_query_tree_nodes"], min(ms[1:]),
            int(bm.astype(np.int64).sum()), int(ids[..., 0].astype(np.int64).sum())), flush=True)
        e.set_option(10, 0); e.close()
