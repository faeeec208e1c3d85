"""Times build variants of the engine (csrc/libvar_*.so, built by hand with extra -D flags) on config 2, whole frame and a 1/8 share — one subprocess per variant, scene arrays built once (GPU box).
Prints ms per frame (min of 4), checksums of bitmap and ids (exactness across variants) and the debug counters."""
import sys, os, subprocess, glob, pickle
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host
CACHE = "/tmp/lib_sweep_c2.pkl"
W, H = 1920, 1080
if len(sys.argv) > 2 and sys.argv[1] == "child":
    import ctypes as C
    engine.LIB_PATH = sys.argv[2]
    a, cam = pickle.load(open(CACHE, "rb"))
    si = wire.default_scene_info(W, H, nb_ray_iterations=3)
    parts = [(0, 1)] + ([(0, 8)] if len(sys.argv) > 3 else [])
    for rank, world in parts:
        e = engine.Engine(si, rank=rank, world=world)
        if os.environ.get('SOLR_MODE'): e.set_option(6, int(os.environ['SOLR_MODE']))
        e.upload(a, randoms=np.zeros(1920 * 1080, np.float32))
        ms = []
        for it in range(5):
            e.render(si, *cam); e.synchronize(); ms.append(e.last_render_ms())
        bm, ids = e.readback(si)
        cnt = (C.c_ulonglong * 8)()
        e.lib.b200_debug_counters(cnt)
        print("%-24s part %d/%d  ms %.3f  checksum %d %d  counters %s" % (os.path.basename(sys.argv[2]), rank, world, min(ms[1:]),
              int(bm.astype(np.int64).sum()), int(ids[..., 0].astype(np.int64).sum()), list(cnt)[2:8]), flush=True)
        if os.environ.get('SOLR_MODE'): e.set_option(6, 1)
        e.close()
else:
    sc = scenes.config2()
    si = wire.default_scene_info(W, H, nb_ray_iterations=3)
    h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
    pickle.dump((a, (sc.eye, sc.target, sc.angles)), open(CACHE, "wb"))
    for lib in sorted(glob.glob(os.path.join(ROOT, "sol-r_b200", "csrc", "libvar_*.so"))):
        try:
            subprocess.call([sys.executable, __file__, "child", lib] + sys.argv[1:2], timeout=45)
        except subprocess.TimeoutExpired:
            print(os.path.basename(lib), "TIMEOUT", flush=True)
