"""A 1/8 share of config 4 (and config 2) on one GPU — the kernels of one rank of an 8-GPU frame split — under option 8 (queue size up
to which a bounce pass keeps its paths in registers) and option 9 (tile order): ms per iteration."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, engine, host, workloads
for key in sys.argv[1:] or ["config4"]:
    wl = workloads.WORKLOADS[key]
    W, H = wl["size"]
    si = workloads.scene_info(key); si.maxPathTracingIterations = 1 << 30
    sc = wl["scene"]()
    h = host.SceneHost(si, limits=wl["limits"], capacity=wl["capacity"]); sc.replay(h); a = h.arrays(); h.close()
    for world in (1, 8):
        for o8, o9 in ((300, 0),) if os.environ.get('SOLR_MODE') else ((300, 0), (0, 0), (100, 0), (600, 0), (1200, 0), (5000, 0), (300, 1)):
            if world == 1 and (o8, o9) != (300, 0): continue
            e = engine.Engine(si, limits=wl["limits"], rank=0, world=world)
            e.set_option(8, o8); e.set_option(9, o9); e.set_option(6, int(os.environ.get('SOLR_MODE', '1')))
            e.upload(a, randoms=np.zeros(max(W * H, 1920 * 1080), np.float32))
            per = {}
            for rep in range(3):
                for it in wl["iterations"]:
                    si.pathTracingIteration = it
                    e.render(si, sc.eye, sc.target, sc.angles); e.synchronize()
                    per.setdefault(it, []).append(e.last_render_ms())
            ms = {it: min(v[1:]) for it, v in per.items()}
            print("%s share 1/%d option8 %5d option9 %d  mean %.3f ms  %s" % (key, world, o8, o9, sum(ms.values()) / len(ms), {k: round(v, 3) for k, v in ms.items()}), flush=True)
            e.set_option(8, 300); e.set_option(9, 0); e.set_option(6, 1); e.close()
