"""Profiling driver for the staged pipeline: prof_staged.py <mode 0|1|2> <gl> <nit> <frames>"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import wire, scenes, engine, host
mode, gl, nit, frames = (int(v) for v in sys.argv[1:5])
sc = scenes.config2(); W, H = 1920, 1080
si = wire.default_scene_info(W, H, graphics_level=gl, nb_ray_iterations=nit)
h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
world = int(os.environ.get('SOLR_WORLD', '1'))
e = engine.Engine(si, rank=0, world=world) if world > 1 else engine.Engine(si); e.set_option(6, mode)
if os.environ.get('SOLR_OPT7'): e.set_option(7, int(os.environ['SOLR_OPT7']))
e.upload(a, randoms=np.zeros(W * H, np.float32))
for it in range(frames):
    e.render(si, sc.eye, sc.target, sc.angles); e.synchronize()
    print("ms", e.last_render_ms(), e.counters(reset=True))
e.set_option(6, 1); e.set_option(7, 0); e.close()
