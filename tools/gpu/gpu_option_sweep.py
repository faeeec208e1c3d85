"""One engine option (b200_set_option key, values) on configs 2 and 4, whole frame and a 1/8 share: ms per frame and checksums
(GPU box).  usage: gpu_option_sweep.py KEY V1,V2,... [c2] [c4]   e.g. 8 0,150,300,600 (small-queue passes), 9 0,1 (tile order);
SOLR_OPT6=3 ... 10 16,24,32 (node visits per slice, build with SOLR_B200_NVCC_FLAGS=-DWITH_TRACE_SLICE)"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host
KEY = int(sys.argv[1]); VALUES = [int(v) for v in sys.argv[2].split(",")]
for cfg in sys.argv[3:] or ["c2", "c4"]:
    if cfg == "c2":
        sc, (W, H), iters, cap = scenes.config2(), (1920, 1080), [0], None
    else:
        sc, (W, H), iters, cap = scenes.config4(), (3840, 2160), [10, 11, 12, 13], (16_000_000, 4_000_000)
    si = wire.default_scene_info(W, H, nb_ray_iterations=3)
    h = host.SceneHost(si, limits=(W, H), capacity=cap); sc.replay(h); a = h.arrays(); h.close()
    for rank, world in ((0, 1), (0, 8)):
        e = engine.Engine(si, limits=(W, H), rank=rank, world=world)
        e.upload(a, randoms=np.zeros(W * H, np.float32))
        if os.environ.get("SOLR_OPT6") is not None:
            e.set_option(6, int(os.environ["SOLR_OPT6"]))   # driver: 0 single kernel, 1 staged, 2 trace queue, 3 sliced walks (experimental builds)
        for pct in VALUES:
            e.set_option(KEY, pct)
            ms = []
            for rep in range(3):
                for it in iters:
                    si.pathTracingIteration = it
                    e.render(si, sc.eye, sc.target, sc.angles); e.synchronize(); ms.append(e.last_render_ms())
            bm, ids = e.readback(si)
            n = len(iters)
            print("%s part %d/%d option %d = %4d  ms/frame %.3f  checksum %d %d" % (cfg, rank, world, KEY, pct, sum(ms[n:]) / len(ms[n:]),
                  int(bm.astype(np.int64).sum()), int(ids[..., 0].astype(np.int64).sum())), flush=True)
        e.close()
