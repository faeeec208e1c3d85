"""Debug: every shadow ray whose order-independent walk disagrees with the ordered one, with a brute-force blocker list
(needs a library built with -DSOLR_DEBUG_SHADOWCMP)."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host
engine.LIB_PATH = sys.argv[1]
sc = scenes.config2(); W, H = 1920, 1080
si = wire.default_scene_info(W, H, graphics_level=4, nb_ray_iterations=1)
h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
e = engine.Engine(si); e.upload(a, randoms=np.zeros(W * H, np.float32))
e.render(si, sc.eye, sc.target, sc.angles); e.readback(si)
lib = ctypes.CDLL(sys.argv[1])
out = np.zeros((64, 32), np.float32)
n = lib.b200_debug_shadowcmp(out.ctypes.data_as(ctypes.c_void_p))
print("disagreements:", n)
prims = a["primitives"] if isinstance(a, dict) else None
np.set_printoptions(suppress=True, linewidth=200)
for k in range(min(n, 64)):
    r = out[k]
    print("o", r[0:3], "d", r[3:6], "ordered", r[6], "unordered", r[7], "light", int(r[8]), "obj", int(r[9]), "lenOL", r[31], "nblockers", int(r[30]))
    for j in range(int(r[30])):
        b = r[10 + 5 * j: 15 + 5 * j]
        print("    idx %d dist %.4f type %d leafpass %d leafT %.5f" % (int(b[0]), b[1], int(b[2]), int(b[3]), b[4]))
e.close()
