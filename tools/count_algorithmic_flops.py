"""Algorithmic work per frame of every bench workload, counted by the ORACLE in the reference's traversal order on sampled rows
(SURVEY.md 8(d) flop weights: oracle_algorithmic_flops) and scaled to the frame -> profiles/algorithmic_flops.json, which bench.py
reads for the roofline of the workloads it does not count live.  CPU only; run here:  python tools/count_algorithmic_flops.py [key ...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from _solr_b200_import import solr_b200  # noqa
from solr_b200 import host, wire, workloads
import oracle

OUT = os.path.join(ROOT, "profiles", "algorithmic_flops.json")
ROWS = {"config2": 27, "config3": 135, "config4": 270, "config5": 270}   # row stride of the sample


def count(key):
    wl = workloads.WORKLOADS[key]
    W, H = wl["size"]
    si = workloads.scene_info(key)
    sc = wl["scene"]()
    h = host.SceneHost(si, limits=wl["limits"], capacity=wl["capacity"])
    sc.replay(h)
    a = h.arrays()
    h.close()
    lim = wl["limits"] or (1920, 1080)
    o = oracle.Oracle(a, W, H, randoms=np.zeros(lim[0] * lim[1], np.float32), random_table_size=lim[0] * lim[1])
    rows = (7, H, ROWS[key])
    n_rows = len(range(*rows))
    per_iteration = {}
    frames = sorted(set([0] + wl["iterations"])) if min(wl["iterations"]) > 0 else wl["iterations"]   # iteration 0 first: later frames read its state
    t0 = time.time()
    for it in frames:
        si.pathTracingIteration = it
        o.render(si, sc.eye, sc.target, sc.angles, rows=rows)
        c = o.counters.as_dict()
        per_iteration[it] = {"gflop_per_frame": o.flops() * H / n_rows / 1e9, "rays_per_frame": c["rays"] * H / n_rows,
                             "box_tests_per_frame": c["box_tests"] * H / n_rows}
    bench_its = wl["iterations"]
    rec = {"workload": wl["name"], "rows_sampled": "%d:%d:%d (%d of %d rows)" % (rows + (n_rows, H)), "primitives": int(a["nbPrimitives"]),
           "boxes": int(a["nbBoxes"]), "per_iteration": {str(k): v for k, v in per_iteration.items()},
           "gflop_per_frame_mean_over_bench_iterations": float(np.mean([per_iteration[i]["gflop_per_frame"] for i in bench_its])),
           "rays_per_frame_mean_over_bench_iterations": float(np.mean([per_iteration[i]["rays_per_frame"] for i in bench_its])),
           "oracle_wall_s": round(time.time() - t0, 1)}
    return rec


if __name__ == "__main__":
    keys = sys.argv[1:] or list(workloads.WORKLOADS)
    try:
        out = json.load(open(OUT))
    except Exception:
        out = {}
    for k in keys:
        out[k] = count(k)
        print(k, json.dumps({x: out[k][x] for x in out[k] if x != "per_iteration"}), flush=True)
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1)
