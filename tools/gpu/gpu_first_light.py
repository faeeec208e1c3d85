"""First GPU bring-up: engine vs oracle (CPU restatement) vs the reference CUDA engine, config 1."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import refh, oracle
from solr_b200 import wire, scenes, engine

def cmp(name, a_bm, a_ids, b_bm, b_ids):
    idm = (a_ids[..., 0] != b_ids[..., 0])
    d = np.abs(a_bm.astype(int) - b_bm.astype(int)).max(axis=-1)
    print("%-28s id mismatches %7d (%.4f%%)  rgb>2: %7d (%.4f%%)  rgb!=: %d  iters!=: %d" % (
        name, idm.sum(), 100.0 * idm.mean(), (d > 2).sum(), 100.0 * (d > 2).mean(), (d > 0).sum(),
        (a_ids[..., 1] != b_ids[..., 1]).sum()))

for cfg, (W, H), nit in (("c1", (1024, 768), 2), ("c2", (1920, 1080), 3)):
    sc = scenes.config1(1000) if cfg == "c1" else scenes.config2()
    si = wire.default_scene_info(W, H, nb_ray_iterations=nit)
    rc = refh.RefScene(si, "cpu")
    t = time.time(); nb = sc.replay(rc); print(cfg, "ref build s", time.time() - t, "boxes", nb, "prims", sc.nb_primitives)
    a = rc.arrays()
    o = oracle.Oracle(a, W, H)
    t = time.time(); obm, oids, opost, k = o.render(si, sc.eye, sc.target, sc.angles); print("oracle s", time.time() - t, k.as_dict())
    print("alg GF", o.flops() / 1e9)
    e = engine.Engine(si)
    e.upload(a, randoms=np.zeros(1920 * 1080, np.float32))
    print(e.scene_stats())
    for it in range(3):
        e.render(si, sc.eye, sc.target, sc.angles)
        bm, ids = e.readback(si)
        print("engine ms", e.last_render_ms(), "counters", e.counters(reset=True))
    cmp("engine vs oracle", bm, ids, obm.copy(), oids.copy())
    if refh.available("cuda"):
        rg = refh.RefScene(si, "cuda")
        sc.replay(rg)
        t = time.time(); gbm, gids, gpost = rg.render(si, sc.eye, sc.target, sc.angles, block=(16, 8)); print("refcuda first s", time.time() - t)
        t = time.time(); gbm, gids, gpost = rg.render(si, sc.eye, sc.target, sc.angles, block=(16, 8)); print("refcuda s (incl h2d/d2h)", time.time() - t)
        cmp("refcuda vs oracle", gbm, gids, obm, oids)
        cmp("engine vs refcuda", bm, ids, gbm, gids)
    e.close()   # before the reference's finalize_scene: it calls cudaDeviceReset() (CudaRayTracer.cu:1530)
    if refh.available("cuda"):
        rg.close()
    rc.close()
    from PIL import Image
    os.makedirs("gpurun_out", exist_ok=True)
    Image.fromarray(bm[::-1]).save("gpurun_out/engine_%s.png" % cfg)
    Image.fromarray(obm[::-1]).save("gpurun_out/oracle_%s.png" % cfg)
