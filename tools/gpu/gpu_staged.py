"""Staged rendering (option key 6) against the single persistent kernel: frames and time, config 1/2 at several levels."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host
for cfg, (W, H), levels in (("c1", (1024, 768), ((4, 2),)), ("c2", (1920, 1080), ((0, 1), (3, 2), (3, 3), (4, 3)))):
    sc = scenes.config1(1000) if cfg == "c1" else scenes.config2()
    for gl, nit in levels:
        si = wire.default_scene_info(W, H, graphics_level=gl, nb_ray_iterations=nit)
        h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
        res = []
        for staged in (0, 1, 2):
            e = engine.Engine(si); e.set_option(6, staged)
            e.upload(a, randoms=np.zeros(1920 * 1080, np.float32))
            ms = []
            for it in range(4):
                e.render(si, sc.eye, sc.target, sc.angles); e.synchronize(); ms.append(e.last_render_ms())
            bm, ids = e.readback(si); post = e.read_post_buffer(si)
            cnt = e.counters(reset=True)
            res.append((bm.copy(), ids.copy(), post.copy(), min(ms[1:]), cnt[0] // 4))
            e.set_option(6, 2); e.close()
        (b0, i0, p0, t0, r0) = res[0]
        for k, (b1, i1, p1, t1, r1) in enumerate(res[1:]):
            print("%s gl %d nit %d: single %.3f ms  staged(%d) %.3f ms  rays %d %d | ids differ %s  post px differ %d  rgb px differ %d" % (
                cfg, gl, nit, t0, k + 1, t1, r0, r1, [(int((i0[..., c] != i1[..., c]).sum())) for c in range(4)],
                int((p0 != p1).any(-1).sum()), int((b0 != b1).any(-1).sum())), flush=True)
