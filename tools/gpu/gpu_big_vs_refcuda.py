"""Large scenes (BASELINE configs 3 and 4 geometry, the reference's own box capacity) engine vs reference CUDA engine."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
import refh
from solr_b200 import wire, scenes, engine
rnd = np.zeros(wire.REF_MAX_BITMAP_SIZE, np.float32)
for cfg in sys.argv[1:] or ["c3"]:
    if cfg == "c3":
        sc = scenes.triangle_mesh(1_000_000); W, H, nit = 960, 540, 5
    else:
        sc = scenes.config4(); W, H, nit = 960, 540, 3
    si = wire.default_scene_info(W, H, nb_ray_iterations=nit)
    t0 = time.time()
    rg = refh.RefScene(si, "cuda"); sc.replay(rg); a = rg.arrays()
    print(cfg, "reference host build %.1fs boxes %d prims %d" % (time.time() - t0, a["nbBoxes"], a["nbPrimitives"]), flush=True)
    e = engine.Engine(si); e.upload(a, randoms=rnd)
    e.render(si, sc.eye, sc.target, sc.angles); e.synchronize(); ms = e.last_render_ms()
    bm, ids = e.readback(si); e.close()
    t0 = time.time()
    gbm, gids, _ = rg.render(si, sc.eye, sc.target, sc.angles, randoms=rnd, block=(16, 8))
    tref = time.time() - t0
    idm = int((ids[..., 0] != gids[..., 0]).sum()); bad = int((np.abs(bm.astype(int) - gbm.astype(int)).max(-1) > 2).sum())
    print("%s %dx%d: engine %.2f ms, reference CUDA %.0f ms (wall) | ids differ %d  rgb>2 %d (%.4f%%)  rgb any %d  hit px %.1f%%" % (
        cfg, W, H, ms, tref * 1e3, idm, bad, 100.0 * bad / (W * H), int((bm != gbm).any(-1).sum()), 100.0 * (ids[..., 0] >= 0).mean()), flush=True)
    rg.close()
