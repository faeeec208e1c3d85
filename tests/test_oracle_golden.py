"""The oracle (oracle/solr_oracle.cpp, the CPU restatement) against the golden vectors the REFERENCE
produced (tests/golden/*.npz, made by tests/golden/make_golden.py from the reference's own device code
compiled for the host).  Bit-exact: ids, RGB8 and the float accumulation buffer."""
import os

import numpy as np
import pytest

import golden_scenes as gs
import oracle
from solr_b200 import host

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_oracle(name, threads=2):
    sc, si, eye, target, angles, rnd, frames = gs.case_setup(name)
    h = host.SceneHost(si)
    sc.replay(h)
    a = h.arrays()
    o = oracle.Oracle(a, si.size.x, si.size.y, randoms=rnd, textures=sc.texture_atlas())
    pp = gs.case_post(name)
    for it in frames:
        si.pathTracingIteration = it
        o.render(si, eye, target, angles, post_info=pp, threads=threads)
    h.close()
    return o, a


@pytest.mark.parametrize("name", sorted(gs.CASES))
def test_oracle_matches_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    o, a = run_oracle(name)
    assert a["nbBoxes"] == int(g["nbBoxes"]) and a["nbPrimitives"] == int(g["nbPrimitives"])
    assert np.array_equal(o.ids, g["ids"]), "id buffer differs at %d pixels" % int((o.ids != g["ids"]).any(-1).sum())
    assert np.array_equal(o.post.view(np.uint32), g["post"].view(np.uint32)), "float accumulation buffer differs"
    assert np.array_equal(o.bitmap, g["bitmap"])


def test_oracle_row_sampling_matches_full_frame():
    """bench.py's cpu_baseline renders a bounded sample of rows; sampled rows must equal the full render."""
    name = "spheres_full"
    sc, si, eye, target, angles, rnd, frames = gs.case_setup(name)
    h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
    full = oracle.Oracle(a, si.size.x, si.size.y, randoms=rnd)
    full.render(si, eye, target, angles, threads=2)
    part = oracle.Oracle(a, si.size.x, si.size.y, randoms=rnd)
    _, _, _, k = part.render(si, eye, target, angles, rows=(3, si.size.y, 8), threads=2)
    assert np.array_equal(part.ids[3::8], full.ids[3::8])
    assert np.array_equal(part.bitmap[3::8], full.bitmap[3::8])
    assert k.pixels == len(range(3, si.size.y, 8)) * si.size.x
    assert part.flops(k) > 0


def test_oracle_counts_rays():
    o, _ = run_oracle("spheres_noshading")
    k = o.counters.as_dict()
    assert k["primary_rays"] == gs.W * gs.H and k["rays"] == gs.W * gs.H and k["shadow_rays"] == 0
    o, _ = run_oracle("spheres_full")
    k = o.counters.as_dict()
    assert k["rays"] > k["primary_rays"] and k["shadow_rays"] > 0 and k["box_tests"] > k["rays"]
