"""Where do engine and reference-CUDA pixels differ?  (GPU box; writes diff masks to gpurun_out/)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import refh
from solr_b200 import wire, scenes, engine
from PIL import Image

cfg = sys.argv[1] if len(sys.argv) > 1 else "c1"
gl = int(sys.argv[2]) if len(sys.argv) > 2 else 4
nit = int(sys.argv[3]) if len(sys.argv) > 3 else 2
W, H = (1024, 768) if cfg == "c1" else (1920, 1080)
sc = scenes.config1(1000) if cfg == "c1" else scenes.config2()
si = wire.default_scene_info(W, H, graphics_level=gl, nb_ray_iterations=nit)
rg = refh.RefScene(si, "cuda"); sc.replay(rg); a = rg.arrays()
e = engine.Engine(si); e.upload(a, randoms=np.zeros(1920 * 1080, np.float32))
e.render(si, sc.eye, sc.target, sc.angles); bm, ids = e.readback(si); post = e.read_post_buffer(si)
gbm, gids, gpost = rg.render(si, sc.eye, sc.target, sc.angles, block=(16, 8))
d = np.abs(post[..., :3] - gpost[..., :3]).max(-1)
dd = np.abs(bm.astype(int) - gbm.astype(int)).max(-1)
print("gl", gl, "nit", nit, "float diff>0: %d  >1e-4: %d  >0.01: %d  >0.1: %d ; rgb8 >2: %d" % ((d > 0).sum(), (d > 1e-4).sum(), (d > 0.01).sum(), (d > 0.1).sum(), (dd > 2).sum()))
print("ids.x != %d  ids.y != %d  ids.z != %d  ids.w != %d" % tuple((ids[..., k] != gids[..., k]).sum() for k in range(4)))
print("depth diff max", np.abs(post[..., 3] - gpost[..., 3]).max(), "depth !=", (post[..., 3] != gpost[..., 3]).sum())
big = d > 0.01
print("of big diffs: engine darker %d, engine brighter %d" % ((post[..., :3].sum(-1) < gpost[..., :3].sum(-1))[big].sum(), (post[..., :3].sum(-1) > gpost[..., :3].sum(-1))[big].sum()))
ys, xs = np.nonzero(big)
for k in range(0, min(len(ys), 2000), max(1, len(ys) // 8)):
    y, x = ys[k], xs[k]
    print("px", x, y, "eng", post[y, x, :4], ids[y, x], "ref", gpost[y, x, :4], gids[y, x])
os.makedirs("gpurun_out", exist_ok=True)
Image.fromarray((big * 255).astype(np.uint8)[::-1]).save("gpurun_out/diffmask_%s_gl%d.png" % (cfg, gl))
e.close()
