"""Profiling driver: renders N frames of a config through the engine (and optionally the reference CUDA
engine) — meant to be wrapped by ncu.  Usage: prof_run.py <c1|c2> <frames> [ref]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import refh
from solr_b200 import wire, scenes, engine

cfg = sys.argv[1]; frames = int(sys.argv[2]); with_ref = len(sys.argv) > 3 and sys.argv[3] == "ref"
W, H, nit = (1024, 768, 2) if cfg == "c1" else (1920, 1080, 3)
sc = scenes.config1(1000) if cfg == "c1" else scenes.config2()
si = wire.default_scene_info(W, H, nb_ray_iterations=nit)
rc = refh.RefScene(si, "cpu"); sc.replay(rc); a = rc.arrays()
e = engine.Engine(si)
e.upload(a, randoms=np.zeros(1920 * 1080, np.float32))
for it in range(frames):
    e.render(si, sc.eye, sc.target, sc.angles)
    e.synchronize()
    print("engine ms", e.last_render_ms(), e.counters(reset=True))
e.close()
if with_ref:
    rg = refh.RefScene(si, "cuda"); sc.replay(rg)
    for it in range(2):
        rg.render(si, sc.eye, sc.target, sc.angles, block=(16, 8))
