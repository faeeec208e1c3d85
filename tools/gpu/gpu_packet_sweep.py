"""Packet-mask sweep: time and parity of each walk policy on config 1/2 (GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host

for cfg, (W, H), nit in (("c1", (1024, 768), 2), ("c2", (1920, 1080), 3)):
    sc = scenes.config1(1000) if cfg == "c1" else scenes.config2()
    si = wire.default_scene_info(W, H, nb_ray_iterations=nit)
    h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
    base = None
    for mask, wide, unord, back in ((0, 1, 0, 1), (0, 1, 1, 1), (0, 1, 1, 0)):
        e = engine.Engine(si)
        e.set_option(2, mask); e.set_option(3, wide); e.set_option(4, unord); e.set_option(5, back)
        e.upload(a, randoms=np.zeros(1920 * 1080, np.float32))
        ms = []
        for it in range(4):
            e.render(si, sc.eye, sc.target, sc.angles); e.synchronize(); ms.append(e.last_render_ms())
        bm, ids = e.readback(si)
        cnt = e.counters(reset=True)
        if base is None:
            base = (bm.copy(), ids.copy())
        same = np.array_equal(bm, base[0]) and np.array_equal(ids, base[1])
        dif = (bm != base[0]).any(-1).sum(); difid = (ids[..., 0] != base[1][..., 0]).sum()
        print(cfg, "mask %2d wide %d unordered %d back %d  ms %.3f  rays/frame %d  identical_to_first %s (rgb px differ %d, id px differ %d)" % (mask, wide, unord, back, min(ms[1:]), cnt[0] // 4, same, dif, difid))
        e.set_option(2, 0); e.set_option(3, 1); e.set_option(4, 1); e.set_option(5, 1)
        e.close()
