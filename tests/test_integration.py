"""The drop-in: the reference's own host library (GPUKernel setters, compactBoxes, frame protocol) with
integration/B200Kernel.cpp as its engine host class, linked against libsolr_b200 — built by
oracle/ref_build/Makefile (`make b200`) where /root/reference exists, travels to the GPU box as
oracle/_ref/libsolr_ref_b200.so."""
import os
import subprocess

import numpy as np
import pytest

import golden_scenes as gs
import refh
from solr_b200 import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/solr"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_b200kernel_compiles_against_unmodified_reference_headers(tmp_path):
    out = tmp_path / "B200Kernel.o"
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-fPIC", "-w", "-c", os.path.join(ROOT, "integration", "B200Kernel.cpp"),
                           "-I" + os.path.join(ROOT, "oracle", "ref_build", "stubs"), "-I" + REF, "-I/usr/local/cuda/include",
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "integration"), "-o", str(out)])
    assert out.exists()


@pytest.mark.gpu
@pytest.mark.skipif(not refh.available("b200"), reason="oracle/_ref/libsolr_ref_b200.so did not travel")
@pytest.mark.parametrize("name", ["mixed_full", "molecule_full", "spheres_progressive"])
def test_reference_host_drives_the_engine_unchanged(name):
    """Reference GPUKernel (its setters, its compactBoxes, its render_begin/render_end protocol) + B200Kernel
    == this repo's SceneHost + the same engine, bit for bit."""
    sc, si, eye, target, angles, rnd, frames = gs.case_setup(name)
    h = host.SceneHost(si)
    sc.replay(h)
    h.set_randoms(rnd, si.timestamp)
    h.set_camera(eye, target, angles)
    h.init_buffers()
    for it in frames:
        si.pathTracingIteration = it
        si.maxPathTracingIterations = it + 1
        h.set_scene_info(si)
        h.set_camera(eye, target, angles)
        h.render_begin(0.0)
        h.render_end()
    bm1, ids1 = h.bitmap().copy(), h.primitive_ids().copy()
    h.close()

    sc, si, eye, target, angles, rnd, frames = gs.case_setup(name)
    r = refh.RefScene(si, "b200")
    sc.replay(r)
    for it in frames:
        si.pathTracingIteration = it
        bm2, ids2, _ = r.render(si, eye, target, angles, randoms=rnd, want_post=False)
    r.close()
    assert np.array_equal(bm1, bm2)
    assert np.array_equal(ids1[..., 0], ids2[..., 0])


@pytest.mark.gpu
@pytest.mark.skipif(not refh.available("b200"), reason="oracle/_ref/libsolr_ref_b200.so did not travel")
def test_drop_in_beyond_the_reference_frame_limit():
    """B200Kernel::setLimits(2560, 1440): GPUKernel's own frame / id / random buffers hold 1920x1080, so the frame lives in
    B200Kernel's buffers and is read through getFrame() / getPrimitiveIdAt() — same pixels and ids as SceneHost at that size."""
    W, H = 2560, 1440
    sc, si, eye, target, angles, rnd, frames = gs.case_setup("molecule_full")
    si.size.x, si.size.y = W, H
    table = np.zeros(W * H, np.float32)
    n = min(rnd.shape[0], 1920 * 1080)
    table[:n] = rnd[:n]
    h = host.SceneHost(si, limits=(W, H))
    sc.replay(h)
    h.set_randoms(table, si.timestamp)
    h.set_camera(eye, target, angles)
    h.init_buffers()
    h.set_lazy_ids(False)
    h.render_begin(0.0)
    h.render_end()
    bm1, ids1 = h.bitmap().copy(), h.primitive_ids().copy()
    h.close()
    r = refh.RefScene(si, "b200", limits=(W, H))
    sc.replay(r)
    bm2, ids2, _ = r.render(si, eye, target, angles, randoms=rnd, want_post=False)
    r.close()
    assert (ids1[..., 0] >= 0).sum() > 0.05 * W * H
    assert np.array_equal(bm1, bm2)
    assert np.array_equal(ids1[..., 0], ids2[..., 0])


@pytest.mark.gpu
@pytest.mark.skipif(not refh.available("b200"), reason="oracle/_ref/libsolr_ref_b200.so did not travel")
def test_reference_host_animates_on_the_device():
    """The reference's own GPUKernel with B200Kernel as engine class: the per-frame loop of MoleculeScene.cpp:75-81 — rotatePrimitives
    + compactBoxes(false) + frame — run by the reference's host code (arrays re-flattened and re-uploaded every step) against
    B200Kernel::rotatePrimitivesOnDevice / translatePrimitivesOnDevice (the arrays moved where they live): same frames; and after
    syncFromDevice the reference container's own flattened arrays are byte-identical to the host-side loop's."""
    moves = [("rotate", ((0.0, 0.0, 0.0), (0.05, 0.2, 0.0))), ("translate", ((15.0, 0.0, -20.0),)), ("rotate", ((0.0, 0.0, 0.0), (0.0, 0.2, 0.1)))]
    out = {}
    for on_device in (False, True):
        sc, si, eye, target, angles, rnd, frames = gs.case_setup("molecule_full")
        r = refh.RefScene(si, "b200")
        sc.replay(r)
        r.render(si, eye, target, angles, randoms=rnd, want_post=False)   # the scene goes up
        shots = []
        for kind, args in moves:
            if on_device:
                assert getattr(r, kind + "_primitives_on_device")(*args), "the step fell back to the host"
            else:
                getattr(r, kind + "_primitives")(*args)
                r.compact_boxes(False)
            bm, ids, _ = r.render(si, eye, target, angles, randoms=rnd, want_post=False)
            shots.append((np.array(bm, copy=True), np.array(ids, copy=True)))
        if on_device:
            r.sync_from_device()
        a = r.arrays()
        r.close()
        out[on_device] = (shots, a)
    for (bm0, id0), (bm1, id1) in zip(out[False][0], out[True][0]):
        assert np.array_equal(bm0, bm1) and np.array_equal(id0, id1)
    for name in ("boxes", "primitives"):
        assert np.array_equal(out[False][1][name], out[True][1][name]), name
    assert not np.array_equal(out[False][0][0][0], out[False][0][1][0])
