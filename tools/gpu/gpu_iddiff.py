import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host
W, H = 1024, 768
sc = scenes.config1(1000)
si = wire.default_scene_info(W, H, graphics_level=0, nb_ray_iterations=1)
h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
out = {}
for mode in (1, 2):
    e = engine.Engine(si); e.set_option(6, mode); e.upload(a, randoms=np.zeros(1920 * 1080, np.float32))
    e.render(si, sc.eye, sc.target, sc.angles); bm, ids = e.readback(si); post = e.read_post_buffer(si)
    out[mode] = (ids.copy(), post.copy()); e.set_option(6, 2); e.close()
i1, p1 = out[1]; i2, p2 = out[2]
ys, xs = np.nonzero(i1[..., 0] != i2[..., 0])
print("differ", len(ys))
for y, x in list(zip(ys, xs))[:30]:
    print((x, y), "mode1 id %d depth %.3f | mode2 id %d depth %.3f" % (i1[y, x, 0], p1[y, x, 3], i2[y, x, 0], p2[y, x, 3]))
