"""Experiment: cost of secondary rays when every lane of a tile has one (all materials reflective) vs the usual third."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host
W, H = 1920, 1080
for allrefl in (0, 1):
    sc = scenes.config2()
    if allrefl:
        sc.mat_f[:-1, 4] = 0.2  # every material but the light
    for gl, nit in ((0, 1), (3, 2), (3, 3), (4, 3)):
        si = wire.default_scene_info(W, H, graphics_level=gl, nb_ray_iterations=nit)
        h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
        e = engine.Engine(si); e.upload(a, randoms=np.zeros(W * H, np.float32))
        ms = []
        for it in range(4):
            e.render(si, sc.eye, sc.target, sc.angles); e.synchronize(); ms.append(e.last_render_ms())
        cnt = e.counters(reset=True)
        print("allrefl %d gl %d nit %d: %.3f ms rays %d" % (allrefl, gl, nit, min(ms[1:]), cnt[0] // 4), flush=True)
        e.close()
