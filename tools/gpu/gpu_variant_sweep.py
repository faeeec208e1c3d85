"""Times kernel build variants (csrc/libvar_*.so) on config 1/2 — one subprocess per variant (GPU box)."""
import sys, os, subprocess, glob
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from _solr_b200_import import solr_b200  # noqa
    from solr_b200 import wire, scenes, engine, host
    engine.LIB_PATH = sys.argv[2]
    masks = [int(m) for m in sys.argv[3].split(",")]
    for cfg, (W, H), nit in (("c1", (1024, 768), 2), ("c2", (1920, 1080), 3)):
        sc = scenes.config1(1000) if cfg == "c1" else scenes.config2()
        si = wire.default_scene_info(W, H, nb_ray_iterations=nit)
        h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
        for mask in masks:
            e = engine.Engine(si); e.set_option(2, mask)
            if os.environ.get('SOLR_OPT5') is not None: e.set_option(5, int(os.environ['SOLR_OPT5']))
            e.upload(a, randoms=np.zeros(1920 * 1080, np.float32))
            ms = []
            for it in range(5):
                e.render(si, sc.eye, sc.target, sc.angles); e.synchronize(); ms.append(e.last_render_ms())
            bm, ids = e.readback(si)
            print("%-28s %s mask %2d  ms %.3f  checksum %d %d" % (os.path.basename(sys.argv[2]), cfg, mask, min(ms[1:]), int(bm.astype(np.int64).sum()), int(ids[..., 0].astype(np.int64).sum())), flush=True)
            e.close()
else:
    masks = sys.argv[1] if len(sys.argv) > 1 else "0,1"
    for lib in sorted(glob.glob(os.path.join(ROOT, "sol-r_b200", "csrc", "libvar_*.so"))):
        subprocess.call([sys.executable, __file__, "child", lib, masks])
