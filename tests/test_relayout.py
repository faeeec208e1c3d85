"""The device box layouts — the ordered SAH BVH over the reference's leaves (default) and the literal layout
(reference hierarchy, single-child chains collapsed) — visit exactly the leaves the reference's flattened
list visits, in the same order, including when the closest-hit distance shrinks during the walk.  Checked on
the host with a scalar walker."""
import numpy as np

import golden_scenes as gs
from solr_b200 import engine, host, scenes, wire

BOX = np.dtype([("lo", "3f4"), ("hi", "3f4"), ("n", "i4"), ("start", "i4"), ("skip", "2i4"), ("pad", "2i4")])


def slab(lo, hi, o, inv, t1):
    f = np.float32
    s = inv < 0
    near = np.where(s, hi, lo); far = np.where(s, lo, hi)
    tn = ((near - o).astype(f) * inv).astype(f); tf = ((far - o).astype(f) * inv).astype(f)
    if tn[0] > tf[1] or tn[1] > tf[0]:
        return False
    tmin = max(tn[0], tn[1]); tmax = min(tf[0], tf[1])
    if tmin > tf[2] or tn[2] > tmax:
        return False
    tmin = max(tmin, tn[2]); tmax = min(tmax, tf[2])
    return bool(tmin < t1 and tmax > 0)


def shrink(t1, start):
    """Stand-in for 'a closer hit was found in this leaf': a deterministic function of the leaf visited."""
    return np.float32(t1 * np.float32(0.93)) if start % 3 == 0 else t1


def walk_reference(boxes, o, inv, t1):
    out, i = [], 0
    while i < len(boxes):
        b = boxes[i]
        if slab(b["lo"], b["hi"], o, inv, t1):
            if b["n"] > 0:
                out.append((int(b["start"]), int(b["n"])))
                t1 = shrink(t1, int(b["start"]))
            i += 1
        else:
            i += int(b["skip"][0])
    return out


def walk_device(packed, o, inv, t1):
    out, i = [], 0
    w = packed.view(np.int32)
    while i < len(packed):
        lo, hi = packed[i, 0:3], packed[i, 4:7]
        w0, w1 = int(w[i, 3]), int(w[i, 7])
        h = slab(lo, hi, o, inv, t1)
        i += 1 if (h or w1 > 0) else w0
        if h and w1 > 0:
            out.append((w0, w1))
            t1 = shrink(t1, w0)
    return out


def check(sc, n_rays=60, seed=1):
    h = host.SceneHost(wire.default_scene_info(64, 48)); sc.replay(h); a = h.arrays(); h.close()
    boxes = np.frombuffer(a["boxes"].tobytes(), dtype=BOX)
    literal = engine.relayout_boxes(a["boxes"], a["nbBoxes"], engine.BOX_LAYOUT_LITERAL)
    packed = engine.relayout_boxes(a["boxes"], a["nbBoxes"])
    assert len(literal) <= len(boxes)
    n_leaves = int((boxes["n"] > 0).sum())
    assert len(packed) == 2 * n_leaves - 1            # a binary tree over the leaves
    w = packed.view(np.int32)
    leaf_rows = w[:, 7] > 0
    assert np.array_equal(w[leaf_rows][:, [3, 7]], np.stack([boxes["start"], boxes["n"]], 1)[boxes["n"] > 0])  # same leaves, same order
    rng = np.random.Generator(np.random.PCG64(seed))
    visited_ref = visited_bvh = 0
    for _ in range(n_rays):
        o = rng.uniform(-9000, 9000, 3).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        d[rng.integers(0, 3)] *= rng.choice([1.0, 0.0, 1e-3])
        inv = np.where(d != 0, np.float32(1) / np.where(d != 0, d, 1), np.float32(1)).astype(np.float32)
        t1 = np.float32(rng.choice([50000.0, 5000.0, 300.0]))
        ref = walk_reference(boxes, o, inv, t1)
        assert ref == walk_device(packed, o, inv, t1)
        assert ref == walk_device(literal, o, inv, t1)
    return len(boxes), len(literal)


def test_collapse_preserves_leaf_visit_order_spheres():
    n_in, n_out = check(scenes.config1(300))
    assert n_out < n_in  # chains exist and were collapsed


def test_collapse_preserves_leaf_visit_order_molecule_and_mesh():
    check(scenes.molecule(cells=1, atoms_per_cell=120), n_rays=40, seed=2)
    check(scenes.triangle_mesh(600), n_rays=40, seed=3)


def test_collapse_of_degenerate_lists():
    assert engine.relayout_boxes(np.zeros(0, np.uint8), 0).shape[0] == 0
    sc, si, *_ = gs.case_setup("textured_skybox")
    n_in, n_out = check(sc, n_rays=20)
    assert n_out >= 1
