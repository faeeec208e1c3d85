import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
import refh
from solr_b200 import wire, scenes, engine
W, H = 1024, 768
sc = scenes.config1(1000)
rnd = np.zeros(wire.REF_MAX_BITMAP_SIZE, np.float32)
gl, nit = int(sys.argv[1]), int(sys.argv[2])
si = wire.default_scene_info(W, H, graphics_level=gl, nb_ray_iterations=nit)
rg = refh.RefScene(si, "cuda"); sc.replay(rg); a = rg.arrays()
res = {}
for unord in (0, 1):
    e = engine.Engine(si); e.set_option(4, unord); e.upload(a, randoms=rnd)
    e.render(si, sc.eye, sc.target, sc.angles); bm, ids = e.readback(si); res[unord] = (bm.copy(), ids.copy()); e.set_option(4, 1); e.close()
gbm, gids, _ = rg.render(si, sc.eye, sc.target, sc.angles, randoms=rnd, block=(16, 8))
for unord in (0, 1):
    bm, ids = res[unord]
    bad = np.abs(bm.astype(int) - gbm.astype(int)).max(-1) > 2
    print("unordered", unord, "rgb>2:", int(bad.sum()), " engine brighter:", int((bad & (bm.astype(int).sum(-1) > gbm.astype(int).sum(-1))).sum()))
bm, ids = res[1]
bad = np.abs(bm.astype(int) - gbm.astype(int)).max(-1) > 2
ys, xs = np.nonzero(bad)
for y, x in list(zip(ys, xs))[:20]:
    print((x, y), "engine", bm[y, x], ids[y, x], "ref", gbm[y, x], gids[y, x])
