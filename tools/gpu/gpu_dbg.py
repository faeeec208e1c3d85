import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host
engine.LIB_PATH = os.path.join(ROOT, "sol-r_b200", "csrc", "libvar_dbg.so")
for cfg, (W, H), gl, nit in (("c2", (1920, 1080), 0, 1), ("c2", (1920, 1080), 4, 1), ("c2", (1920, 1080), 3, 3), ("c2", (1920, 1080), 4, 3)):
    sc = scenes.config2() if cfg == "c2" else scenes.random_spheres(1_000_000, 20000.0, 20.0, 60.0, scenes.SEED + 4, "c4")
    si = wire.default_scene_info(W, H, graphics_level=gl, nb_ray_iterations=nit)
    h = host.SceneHost(si, capacity=(16_000_000, 4_000_000)); sc.replay(h); a = h.arrays(); h.close()
    e = engine.Engine(si); e.upload(a, randoms=np.zeros(1920 * 1080, np.float32))
    e.counters(reset=True)
    e.render(si, sc.eye, sc.target, sc.angles); e.synchronize()
    out = (C.c_ulonglong * 8)()
    e.lib.b200_debug_counters(out)
    v = list(out)
    print("   max visits", v[6], " walks > 200 visits", v[1], " > 1000", v[0] )
    print(cfg, "gl", gl, "nit", nit, "ms %.2f" % e.last_render_ms(), "rays %d | walks %d overflow %d cands %d | wide-node visits %d (%.1f/walk) leaf visits %d (%.1f/walk) prim tests %d (%.1f/walk)" % (
        v[0], v[2], v[3], v[4], v[5], v[5] / max(v[2], 1), v[6], v[6] / max(v[2], 1), v[7], v[7] / max(v[2], 1)), flush=True)
    e.close()
