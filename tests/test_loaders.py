"""The loaders' hot spot: the bond search of PDBReader.cpp:616-660 from a uniform grid (csrc/loaders.cpp) must return
exactly the pairs, in exactly the order, of the reference's loop over every pair of atoms — restated literally here."""
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
