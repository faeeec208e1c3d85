import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host
sc = scenes.config2(); W, H = 1920, 1080
gl, nit = int(sys.argv[1]), int(sys.argv[2])
if len(sys.argv) > 3: engine.LIB_PATH = sys.argv[3]
si = wire.default_scene_info(W, H, graphics_level=gl, nb_ray_iterations=nit)
h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
res = []
for unord in (0, 1):
    e = engine.Engine(si); e.set_option(4, unord); e.upload(a, randoms=np.zeros(1920 * 1080, np.float32))
    e.render(si, sc.eye, sc.target, sc.angles); bm, ids = e.readback(si); post = e.read_post_buffer(si)
    res.append((bm.copy(), ids.copy(), post.copy(), e.counters(reset=True)))
    e.set_option(4, 1); e.close()
(b0, i0, p0, c0), (b1, i1, p1, c1) = res
print("gl", gl, "nit", nit, "rays", c0[0], c1[0])
for k in range(4):
    print("ids[%d] differ: %d" % (k, (i0[..., k] != i1[..., k]).sum()))
d = (p0[..., :3] != p1[..., :3]).any(-1)
print("post differ px:", d.sum(), " rgb differ px:", (b0 != b1).any(-1).sum())
ys, xs = np.nonzero(d)
for k in range(min(8, len(ys))):
    y, x = ys[k], xs[k]
    print((x, y), "ordered", i0[y, x], p0[y, x, :3], "unordered", i1[y, x], p1[y, x, :3])
