"""Host-side checks (no GPU) of the trees the order-independent walks use (engine.cu buildWalkTrees):
structure — every primitive is a leaf of the main tree exactly once, child boxes contain what hangs below them, the
point-query tree lists cylinders/cones only, with boxes that contain the primitive's own box — and function: a scalar
walk of the records finds every primitive whose box a ray crosses, for any ray (that is all the trees have to promise:
hits are re-checked against the reference's own leaf box and primitive test)."""
import numpy as np
import pytest

import golden_scenes as gs
from solr_b200 import engine, host, scenes, wire

EMPTY = -(1 << 31)


def build(scene):
    si = wire.default_scene_info(96, 72)
    h = host.SceneHost(si)
    scene.replay(h)
    a = h.arrays()
    h.close()
    nodes, prim_leaf, nb_main, nb_ext = engine.build_walk_trees(a)
    return a, nodes, prim_leaf, nb_main, nb_ext


def children(nodes, k):
    """[(lo[3], hi[3], ref)] of node k (one 128-byte record: rows lo.x lo.y lo.z hi.x hi.y hi.z refs)."""
    rec = nodes[k]
    refs = rec[6].view(np.int32)
    out = []
    for c in range(4):
        if refs[c] != EMPTY:
            out.append((rec[0:3, c].copy(), rec[3:6, c].copy(), int(refs[c])))
    return out


def prim_tight_box(P, i):
    p = P[i]
    t = int(p.view(np.int32)[21])
    p0, p1, p2, size = p[0:3], p[3:6], p[6:9], p[18:21]
    if t == wire.PT_TRIANGLE:
        return np.minimum(np.minimum(p0, p1), p2), np.maximum(np.maximum(p0, p1), p2)
    if t in (wire.PT_CYLINDER, wire.PT_CONE):
        return np.minimum(p0, p1) - abs(size[0]), np.maximum(p0, p1) + abs(size[0])
    if t in (wire.PT_SPHERE, wire.PT_ENVIRONMENT):
        return p0 - abs(size[0]), p0 + abs(size[0])
    return p0 - np.abs(size), p0 + np.abs(size)


@pytest.mark.parametrize("maker", [gs.spheres_scene, gs.molecule_scene, gs.mixed_scene, gs.mesh_scene])
def test_walk_tree_structure(maker):
    a, nodes, prim_leaf, nb_main, nb_ext = build(maker())
    m = a["nbPrimitives"]
    P = a["primitives"].view(np.float32).reshape(-1, 32)
    assert nb_main >= 1 and nodes.shape[0] == nb_main + nb_ext
    seen = np.zeros(m, np.int32)

    def visit(k, lo_parent, hi_parent, ext):
        for lo, hi, ref in children(nodes, k):
            if lo_parent is not None:
                assert np.all(lo >= lo_parent - 1e-3) and np.all(hi <= hi_parent + 1e-3), "child box sticks out of its parent"
            if ref >= 0:
                assert (ref >= nb_main) == ext
                visit(ref, lo, hi, ext)
            else:
                idx = (~ref) & 0x3FFFFFFF
                assert bool((~ref) & 0x40000000) == ext and 0 <= idx < m
                tlo, thi = prim_tight_box(P, idx)
                if ext:
                    assert int(P[idx].view(np.int32)[21]) in (wire.PT_CYLINDER, wire.PT_CONE)
                else:
                    seen[idx] += 1
                    assert np.all(lo <= tlo) and np.all(hi >= thi), "leaf box does not contain its primitive"

    visit(0, None, None, False)
    assert np.all(seen == 1), "every primitive exactly once in the main tree"
    if nb_ext:
        visit(nb_main, None, None, True)
    # primLeaf: the reference leaf (array order) holding each primitive
    boxes = a["boxes"].view(np.int32).reshape(-1, 12)
    leaves = boxes[boxes[:, 6] > 0]
    for i in range(0, m, max(1, m // 50)):
        l = leaves[prim_leaf[i]]
        assert l[7] <= i < l[7] + l[6]


def test_scalar_walk_finds_every_box_crossing():
    a, nodes, prim_leaf, nb_main, nb_ext = build(gs.molecule_scene())
    m = a["nbPrimitives"]
    P = a["primitives"].view(np.float32).reshape(-1, 32)
    tight = [prim_tight_box(P, i) for i in range(m)]
    rng = np.random.Generator(np.random.PCG64(7))
    lo_all = np.min([t[0] for t in tight], 0); hi_all = np.max([t[1] for t in tight], 0)
    for _ in range(60):
        o = rng.uniform(lo_all - 200, hi_all + 200).astype(np.float64)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        inv = 1.0 / d

        def crosses(lo, hi):
            t1 = (lo - o) * inv; t2 = (hi - o) * inv
            tn = np.minimum(t1, t2).max(); tf = np.maximum(t1, t2).min()
            return tn <= tf and tf > 0

        found = set()
        stack = [0]
        while stack:
            k = stack.pop()
            for lo, hi, ref in children(nodes, k):
                if crosses(lo.astype(np.float64), hi.astype(np.float64)):
                    if ref >= 0:
                        stack.append(ref)
                    else:
                        found.add(~ref)
        expect = {i for i in range(m) if crosses(tight[i][0].astype(np.float64), tight[i][1].astype(np.float64))}
        assert expect <= found
