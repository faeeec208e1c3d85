"""Where do frames differ between walk trees with / without split cylinders and the GPU-built trees? (molecule, 640x360)"""
import sys, os, subprocess, pickle
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host
import golden_scenes as gs
W, H = 640, 360
if len(sys.argv) > 1 and sys.argv[1] == "child":
    engine.LIB_PATH = sys.argv[2]
    a, cam, rnd = pickle.load(open("/tmp/splitdiff.pkl", "rb"))
    si = wire.default_scene_info(W, H, graphics_level=wire.GL_FULL, nb_ray_iterations=3)
    e = engine.Engine(si); e.set_option(10, int(sys.argv[3]))
    e.upload(a, randoms=rnd); e.render(si, *cam)
    bm, ids = e.readback(si); post = e.read_post_buffer(si)
    pickle.dump((bm, ids, post), open(sys.argv[4], "wb")); e.set_option(10, 0); e.close()
else:
    sc = scenes.molecule(cells=3)
    si = wire.default_scene_info(W, H, graphics_level=wire.GL_FULL, nb_ray_iterations=3)
    h = host.SceneHost(si); sc.replay(h); h.rotate_primitives((0.0, 0.0, 0.0), (0.1, 0.25, 0.05)); h.compact_boxes(False); a = dict(h.arrays()); h.close()
    pickle.dump((a, (sc.eye, sc.target, sc.angles), gs.randoms(47)), open("/tmp/splitdiff.pkl", "wb"))
    res = {}
    for name, lib, opt in (("split8", "libvar_split8.so", 0), ("split0", "libvar_split0.so", 0), ("gpu", "libvar_split8.so", 1)):
        out = "/tmp/splitdiff_%s.pkl" % name
        subprocess.check_call([sys.executable, __file__, "child", os.path.join(ROOT, "sol-r_b200", "csrc", lib), str(opt), out])
        res[name] = pickle.load(open(out, "rb"))
    for x, y in (("split8", "split0"), ("split0", "gpu"), ("split8", "gpu")):
        d = np.abs(res[x][2][..., :3] - res[y][2][..., :3]).max(-1)
        ys, xs = np.nonzero(d)
        print(x, "vs", y, "ids differing", int((res[x][1] != res[y][1]).any(-1).sum()), "post pixels differing", len(ys), "max", float(d.max()),
              "bitmap pixels differing", int((res[x][0] != res[y][0]).any(-1).sum()), "first", list(zip(ys[:6].tolist(), xs[:6].tolist())))
        for yy, xx in list(zip(ys[:4].tolist(), xs[:4].tolist())):
            print("   ", (yy, xx), res[x][2][yy, xx], res[y][2][yy, xx], res[x][1][yy, xx], res[y][1][yy, xx])
