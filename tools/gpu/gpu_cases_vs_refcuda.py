"""Every golden case (tests/golden_scenes.py) at 4x the golden resolution: engine vs the reference CUDA engine on this GPU."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
import refh, golden_scenes as gs
from solr_b200 import wire, scenes, engine
S = 4
for name in sorted(gs.CASES):
    sc, si, eye, target, angles, rnd, frames = gs.case_setup(name)
    pp = gs.case_post(name)
    si.size.x *= S; si.size.y *= S
    W, H = si.size.x, si.size.y
    rg = refh.RefScene(si, "cuda"); sc.replay(rg); a = rg.arrays()
    atlas = sc.texture_atlas()
    e = engine.Engine(si)
    tex = None
    if atlas is not None:
        infos = (wire.TextureInfo * 1)(); infos[0].buffer = atlas.ctypes.data; infos[0].offset = 0
        infos[0].size = wire.Int3(int(atlas.shape[0]), 1, 1); tex = (infos, 1)
    e.upload(a, randoms=rnd, textures=tex)
    for it in frames:
        si.pathTracingIteration = it
        e.render(si, eye, target, angles, post_info=pp)
    bm, ids = e.readback(si); e.close()
    for it in frames:
        si.pathTracingIteration = it
        gbm, gids, _ = rg.render(si, eye, target, angles, randoms=rnd, post_info=pp, block=(16, 8))
    idm = int((ids[..., 0] != gids[..., 0]).sum())
    bad = int((np.abs(bm.astype(int) - gbm.astype(int)).max(-1) > 2).sum())
    anyd = int((bm != gbm).any(-1).sum())
    print("%-28s %dx%d: ids differ %5d  rgb>2 %6d (%.3f%%)  rgb any %6d" % (name, W, H, idm, bad, 100.0 * bad / (W * H), anyd), flush=True)
    rg.close()
