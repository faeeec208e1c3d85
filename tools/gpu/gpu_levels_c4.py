import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sc = scenes.random_spheres(n, 20000.0, 20.0, 60.0, scenes.SEED + 4, "c4")
W, H = 1920, 1080
si0 = wire.default_scene_info(W, H)
h = host.SceneHost(si0, capacity=(16_000_000, 4_000_000)); sc.replay(h); a = h.arrays(); h.close()
for gl, nit, it in ((0, 1, 0), (4, 1, 0), (3, 3, 0), (4, 3, 0), (4, 3, 10)):
    for unord in (0, 1):
        si = wire.default_scene_info(W, H, graphics_level=gl, nb_ray_iterations=nit)
        si.pathTracingIteration = it; si.maxPathTracingIterations = 100
        e = engine.Engine(si); e.set_option(4, unord); e.upload(a, randoms=np.zeros(1920 * 1080, np.float32))
        ms = []
        for k in range(3):
            e.render(si, sc.eye, sc.target, sc.angles); e.synchronize(); ms.append(e.last_render_ms())
        cnt = e.counters(reset=True)
        print("n %d gl %d nit %d iter %d unordered %d: %.3f ms  rays %d -> %.0f Mrays/s" % (n, gl, nit, it, unord, min(ms[1:]), cnt[0] // 3, cnt[0] / 3 / min(ms[1:]) / 1e3), flush=True)
        e.set_option(4, 1); e.close()
