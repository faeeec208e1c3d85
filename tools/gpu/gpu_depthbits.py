"""Engine vs reference CUDA engine: first-hit depth (colorInfo.w) and colour accumulation, bit for bit, primary rays only."""
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
for gl in (0, 2, 4):
    si = wire.default_scene_info(W, H, graphics_level=gl, nb_ray_iterations=1)
    rg = refh.RefScene(si, "cuda"); sc.replay(rg); a = rg.arrays()
    e = engine.Engine(si); e.upload(a, randoms=rnd)
    e.render(si, sc.eye, sc.target, sc.angles); bm, ids = e.readback(si); post = e.read_post_buffer(si).copy(); e.close()
    gbm, gids, gpost = rg.render(si, sc.eye, sc.target, sc.angles, randoms=rnd, block=(16, 8))
    same_id = ids[..., 0] == gids[..., 0]
    d_bits = (post[..., 3].view(np.uint32) != gpost[..., 3].view(np.uint32)) & same_id
    c_bits = (post[..., :3].view(np.uint32) != gpost[..., :3].view(np.uint32)).any(-1) & same_id
    dd = np.abs(post[..., 3] - gpost[..., 3])[same_id]
    print("gl %d: ids differ %d | depth bits differ %d px (max abs %.6f) | colour bits differ %d px | rgb8 differ %d px" % (
        gl, int((~same_id).sum()), int(d_bits.sum()), float(dd.max()), int(c_bits.sum()), int((bm != gbm).any(-1).sum())), flush=True)

