"""Engine (each staging mode) against the reference CUDA engine built for sm_100 (oracle/_ref): id and rgb mismatch."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
import refh
from solr_b200 import wire, scenes, engine
if len(sys.argv) > 1: engine.LIB_PATH = sys.argv[1]
for cfg in ("config1", "molecule"):
    W, H = (1024, 768) if cfg == "config1" else (960, 540)
    sc = scenes.config1(1000) if cfg == "config1" else scenes.molecule(cells=3)
    rnd = np.zeros(wire.REF_MAX_BITMAP_SIZE, np.float32)
    for gl, nit in ((wire.GL_PHONG_BLINN, 1), (wire.GL_FULL, 1), (wire.GL_REFLECTIONS, 2), (wire.GL_REFLECTIONS, 3), (wire.GL_FULL, 3)):
        si = wire.default_scene_info(W, H, graphics_level=gl, nb_ray_iterations=nit)
        rg = refh.RefScene(si, "cuda"); sc.replay(rg); a = rg.arrays()
        frames = {}
        for mode in (1,):
            e = engine.Engine(si); e.set_option(6, mode); e.upload(a, randoms=rnd)
            e.render(si, sc.eye, sc.target, sc.angles); bm, ids = e.readback(si)
            frames[mode] = (bm.copy(), ids.copy()); e.set_option(6, 1); e.close()
        gbm, gids, _ = rg.render(si, sc.eye, sc.target, sc.angles, randoms=rnd, block=(16, 8))
        for mode, (bm, ids) in frames.items():
            idm = int((ids[..., 0] != gids[..., 0]).sum())
            bad = int((np.abs(bm.astype(int) - gbm.astype(int)).max(-1) > 2).sum())
            anyd = int((bm != gbm).any(-1).sum())
            print("%s gl %d nit %d mode %d: ids differ %d px (%.4f%%)  rgb>2 %d px (%.3f%%)  rgb any %d px  iter differ %d" % (cfg, gl, nit, mode, idm, 100.0 * idm / (W * H), bad, 100.0 * bad / (W * H), anyd, int((ids[..., 1] != gids[..., 1]).sum())), flush=True)
