"""Time per graphics level / bounce depth on config 2 (how the frame time splits over ray classes)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host
sc = scenes.config2(); W, H = 1920, 1080
si0 = wire.default_scene_info(W, H)
h = host.SceneHost(si0); sc.replay(h); a = h.arrays(); h.close()
for gl, nit in ((0, 1), (2, 1), (4, 1), (3, 3), (4, 3)):
    for mask in (0, 100):
        si = wire.default_scene_info(W, H, graphics_level=gl, nb_ray_iterations=nit)
        e = engine.Engine(si); e.set_option(2, 0); e.set_option(4, 1 if mask == 100 else 0); e.upload(a, randoms=np.zeros(1920 * 1080, np.float32))
        ms = []
        for it in range(4):
            e.render(si, sc.eye, sc.target, sc.angles); e.synchronize(); ms.append(e.last_render_ms())
        cnt = e.counters(reset=True)
        print("gl %d nit %d mask %d: %.3f ms  rays %d  -> %.0f Mrays/s" % (gl, nit, mask, min(ms[1:]), cnt[0] // 4, cnt[0] / 4 / min(ms[1:]) / 1e3))
        e.set_option(4, 1); e.close()
