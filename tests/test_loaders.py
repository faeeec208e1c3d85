"""The loaders' hot spot: the bond search of PDBReader.cpp:616-660 from a uniform grid (csrc/loaders.cpp) must return
exactly the pairs, in exactly the order, of the reference's loop over every pair of atoms — restated literally here."""
import ctypes
import time

import numpy as np
import pytest

from solr_b200 import host


def literal(xyz, processed, backbone, stick, backbone_geometry):
    """PDBReader.cpp:616-660: for every atom, every other atom in map order; float arithmetic as the reference's."""
    n = xyz.shape[0]
    first, partners = [0], []
    for i in range(n):
        d = xyz[i] - xyz                                   # float32
        dist = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(np.float32)
        limit = np.where(backbone_geometry & (backbone != 0), np.float32(stick * 2.0), np.float32(stick)).astype(np.float32)
        ok = (np.arange(n) != i) & (processed < 2) & (backbone == backbone[i]) & (dist < limit)
        partners.extend(np.nonzero(ok)[0].tolist())
        first.append(len(partners))
    return np.array(first, np.int32), np.array(partners, np.int32)


@pytest.mark.parametrize("backbone_geometry", [False, True])
@pytest.mark.parametrize("seed,n,extent", [(1, 1, 5.0), (2, 400, 6.0), (3, 3000, 25.0), (4, 2000, 4000.0)])
def test_find_bonds_equals_the_loop_over_all_pairs(seed, n, extent, backbone_geometry):
    rng = np.random.Generator(np.random.PCG64(seed))
    xyz = rng.uniform(-extent, extent, size=(n, 3)).astype(np.float32)
    if n > 10:
        xyz[5] = xyz[4]                                    # coincident atoms: distance 0
        xyz[7] = xyz[6] + np.float32(1.7)                  # exactly on the limit along one axis
    processed = rng.integers(0, 3, size=n).astype(np.int32)
    backbone = rng.integers(0, 2, size=n).astype(np.uint8)
    f0, p0 = literal(xyz, processed, backbone, 1.7, backbone_geometry)
    f1, p1 = host.find_bonds(xyz, processed, backbone, 1.7, backbone_geometry)
    assert np.array_equal(f0, f1)
    assert np.array_equal(p0, p1)
    if n >= 400 and extent < 100:
        assert len(p0) > 0


def test_find_bonds_at_the_size_of_config_2():
    """100 k atoms: the reference's loop is 10^10 distance tests; the grid takes a fraction of a second.  Spot-checked against the
    literal loop on a sample of atoms."""
    rng = np.random.Generator(np.random.PCG64(9))
    n = 100_000
    xyz = rng.uniform(-40.0, 40.0, size=(n, 3)).astype(np.float32)
    processed = np.zeros(n, np.int32)
    backbone = rng.integers(0, 2, size=n).astype(np.uint8)
    t0 = time.perf_counter()
    first, partners = host.find_bonds(xyz, processed, backbone, 1.7, False)
    dt = time.perf_counter() - t0
    assert dt < 5.0
    for i in rng.integers(0, n, size=40):
        d = xyz[i] - xyz
        dist = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(np.float32)
        ok = (np.arange(n) != i) & (backbone == backbone[i]) & (dist < np.float32(1.7))
        assert np.array_equal(np.nonzero(ok)[0].astype(np.int32), partners[first[i]:first[i + 1]])


# ---- OBJ reader, first pass (OBJReader.cpp:440-563) ------------------------------------------------------------------------------
_atof = ctypes.CDLL(None).atof
_atof.restype = ctypes.c_double
_atof.argtypes = [ctypes.c_char_p]


def literal_obj_vertex_pass(text):
    """The reference's loop, statement by statement: getline, carriage returns removed, lines longer than one character that start
    with 'v'; a blank after a non-blank closes an item, atof on what was collected, the last item taken at the end of the line;
    maps keyed from 1 (here: lists)."""
    vertices, normals, tex = [], [], []
    box = [np.float32(100000.0)] * 3 + [np.float32(-100000.0)] * 3
    for raw in text.split("\n"):
        line = raw.replace("\r", "")
        if len(line) <= 1 or line[0] != "v":
            continue
        v = [np.float32(0.0)] * 3
        value, item, previous, i = "", 0, line[0], 1

        def assign(it, value):
            if 1 <= it <= 3:
                v[it - 1] = np.float32(_atof(value.encode("latin-1")))
        while i < len(line) and item < 4:
            if line[i] == " " and previous != " ":
                assign(item, value)
                item += 1
                value = ""
            else:
                value += line[i]
            previous = line[i]
            i += 1
        if len(value) != 0:
            assign(item, value)
        if line[1] == "n":
            normals.append((v[0], v[1], -v[2]))
        elif line[1] == "t":
            x, y = v[0], v[1]
            if x < 0:
                x = np.float32(abs(x) - np.float32(int(abs(x))))
            if y < 0:
                y = np.float32(abs(y) - np.float32(int(abs(y))))
            tex.append((x, y))
        elif line[1] == " ":
            p = (v[0], v[1], -v[2])
            vertices.append(p)
            for a in range(3):
                box[a] = p[a] if p[a] < box[a] else box[a]
                box[3 + a] = p[a] if p[a] > box[3 + a] else box[3 + a]
    f = lambda rows, w: np.array(rows, np.float32).reshape(-1, w)
    return f(vertices, 3), f(normals, 3), f(tex, 2), np.array(box, np.float32)


def _obj_text(rng, n):
    """Lines as exporters write them, and as they should not: several blanks, tabs, carriage returns, a fourth component, missing
    components, exponents, garbage after a number, comments, faces, keywords that merely start with 'v', no newline at the end."""
    out = ["# a comment", "mtllib scene.mtl", "g SoL_R_light", "vertex 1 2 3", "v", "vp 0.5 0.5"]
    for k in range(n):
        x, y, z = rng.uniform(-5000, 5000, 3)
        kind = rng.integers(0, 12)
        if kind == 0:
            out.append("v  %.6f   %.6f %.6f" % (x, y, z))
        elif kind == 1:
            out.append("v %.4f %.4f %.4f 1.0\r" % (x, y, z))
        elif kind == 2:
            out.append("vn %.5f %.5f %.5f" % tuple(rng.uniform(-1, 1, 3)))
        elif kind == 3:
            out.append("vt %.5f %.5f" % tuple(rng.uniform(-3, 3, 2)))
        elif kind == 4:
            out.append("vt %.5f %.5f 0.0 " % tuple(rng.uniform(-3, 3, 2)))
        elif kind == 5:
            out.append("v %e %e %e" % (x, y, z))
        elif kind == 6:
            out.append("v\t%.3f %.3f %.3f" % (x, y, z))          # a tab is not a separator: the line has no "v " shape
        elif kind == 7:
            out.append("v %.3f %.3f" % (x, y))                    # z stays 0, then negated
        elif kind == 8:
            out.append("v %.3fabc %.3f, %.3f;" % (x, y, z))       # atof stops at the first foreign character
        elif kind == 9:
            out.append("f %d/%d/%d %d/%d/%d %d/%d/%d" % tuple(rng.integers(1, 50, 9)))
        elif kind == 10:
            out.append("v %.6f %.6f %.6f " % (x, y, z))           # trailing blank
        else:
            out.append("v %d %d %d" % (int(x), int(y), int(z)))
    return "\n".join(out)


@pytest.mark.parametrize("seed,n", [(1, 0), (2, 50), (3, 3000)])
def test_obj_vertex_pass_equals_the_reference_loop(seed, n):
    text = _obj_text(np.random.Generator(np.random.PCG64(seed)), n)
    v0, n0, t0, b0 = literal_obj_vertex_pass(text)
    v1, n1, t1, b1 = host.obj_vertex_pass(text)
    for a, b in ((v0, v1), (n0, n1), (t0, t1), (b0, b1)):
        assert a.shape == b.shape
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))   # bit for bit, -0.0 included
    if n >= 50:
        assert len(v0) > 0 and len(n0) > 0 and len(t0) > 0


def test_obj_vertex_pass_at_the_size_of_config_3():
    """500 k vertices + 500 k normals (the 1 M-triangle mesh of config 3): one pass over 40 MB of text."""
    rng = np.random.Generator(np.random.PCG64(11))
    n = 500_000
    xyz = rng.uniform(-5000, 5000, size=(n, 3))
    nrm = rng.uniform(-1, 1, size=(n, 3))
    text = "\n".join(["v %.6f %.6f %.6f" % tuple(r) for r in xyz] + ["vn %.6f %.6f %.6f" % tuple(r) for r in nrm]) + "\n"
    t0 = time.perf_counter()
    v, nn, t, box = host.obj_vertex_pass(text)
    dt = time.perf_counter() - t0
    assert v.shape == (n, 3) and nn.shape == (n, 3) and t.shape == (0, 2)
    assert dt < 10.0
    expect = np.array([[float("%.6f" % c) for c in r] for r in xyz[:200]], np.float64)
    assert np.array_equal(v[:200], (expect * [1, 1, -1]).astype(np.float32))
    assert np.array_equal(box, np.concatenate([v.min(0), v.max(0)]))
