"""Streamed output (include/solr_b200.h, option key 12): frames whose reader reads every frame into registered host buffers are
written there by the ray kernels themselves, tile by tile.  The bytes must be the ones the copy delivers (the reference's
render_end, CudaKernel.cpp:304-313), on every frame of a progressive sequence — deepening passes rewrite only some pixels
(CudaRayTracer.cu:454-458), accumulation passes all of them — for both staged cameras, and the frames the mechanism does not
apply to must fall back to the copy silently."""
import numpy as np
import pytest

import golden_scenes as gs
from solr_b200 import engine, host, wire

pytestmark = pytest.mark.gpu


def _arrays(name):
    sc, si, eye, target, angles, rnd, frames = gs.case_setup(name)
    h = host.SceneHost(si); sc.replay(h); a = h.arrays(); h.close()
    return a, si, eye, target, angles, rnd, frames


def _sequence(name, streamed, frames=None, lazy_ids=False, post=None):
    """Renders the case's frames, reading every frame back into ONE pair of buffers (registered when `streamed`); returns the
    per-frame copies of what the reader saw, and how many frames were streamed."""
    a, si, eye, target, angles, rnd, case_frames = _arrays(name)
    frames = case_frames if frames is None else frames
    e = engine.Engine(si)
    e.set_option(12, 1 if streamed else 0)
    e.upload(a, randoms=rnd)
    H, W = si.size.y, si.size.x
    bm = np.zeros((H, W, 3), np.uint8)
    ids = np.zeros((H, W, 4), np.int32)
    if streamed:
        assert e.register_host(bm) == 0 and e.register_host(ids) == 0
    n0 = e.frames_streamed()
    seen = []
    for it in frames:
        si.pathTracingIteration = it
        e.render(si, eye, target, angles, post_info=post)
        if lazy_ids:
            e.lib.b200_d2h_bitmap(e.OCC, si, bm.ctypes.data, None)
            e.check()
            seen.append((bm.copy(), None))
        else:
            e.readback(si, bitmap=bm, ids=ids)
            seen.append((bm.copy(), ids.copy()))
    n = e.frames_streamed() - n0
    if lazy_ids:
        e.lib.b200_d2h_bitmap(e.OCC, si, None, ids.ctypes.data)   # on demand, after the last frame
        e.check()
        seen.append((None, ids.copy()))
    if streamed:
        e.unregister_host(bm); e.unregister_host(ids)
    e.set_option(12, 1)
    e.close()
    return seen, n


def _same(a, b):
    assert len(a) == len(b)
    for k, ((bm1, id1), (bm2, id2)) in enumerate(zip(a, b)):
        if bm1 is not None:
            assert np.array_equal(bm1, bm2), "frame %d: bitmap differs in %d bytes" % (k, int(np.count_nonzero(bm1 != bm2)))
        if id1 is not None:
            assert np.array_equal(id1, id2), "frame %d: ids differ in %d words" % (k, int(np.count_nonzero(id1 != id2)))


@pytest.mark.parametrize("name", ["spheres_progressive", "spheres_anaglyph", "mixed_full", "spheres_rotated"])
def test_streamed_frames_equal_copied_frames(name):
    frames = {"spheres_anaglyph": [0, 1, 2, 10, 11, 12], "mixed_full": [0, 0, 0], "spheres_rotated": [0, 0]}.get(name)
    copied, n_c = _sequence(name, False, frames)
    streamed, n_s = _sequence(name, True, frames)
    assert n_c == 0
    assert n_s == len(copied) - 1   # the first frame is copied (it arms the buffers), every later one is streamed
    _same(copied, streamed)


def test_pixels_alone_are_copied():
    """The drop-in's lazy-id protocol: only the bitmap is read every frame, the ids once at the end.  Six megabytes of pixels are
    cheaper to copy than to count tiles for (measured: profiles/r02_history.md), so nothing is streamed — same bytes either way."""
    copied, _ = _sequence("spheres_progressive", False, lazy_ids=True)
    streamed, n = _sequence("spheres_progressive", True, lazy_ids=True)
    assert n == 0
    _same(copied, streamed)


def test_frames_the_mechanism_does_not_cover_are_copied():
    """A post-processing effect rewrites the frame after the ray kernels; the single-kernel cameras do not count tiles."""
    pp = gs.case_post("spheres_pp_dof")
    for name, post in (("spheres_pp_dof", pp), ("spheres_aa", None)):
        copied, _ = _sequence(name, False, [0, 0, 0], post=post)
        streamed, n = _sequence(name, True, [0, 0, 0], post=post)
        assert n == 0
        _same(copied, streamed)


def test_a_frame_nobody_read_is_not_streamed_and_the_next_read_is_whole():
    """render, read, render, render, read: the third frame has no reader in between, so it is not streamed and the read after it
    copies; unregistered buffers are never written by kernels."""
    a, si, eye, target, angles, rnd, _ = _arrays("spheres_full")
    e = engine.Engine(si)
    e.upload(a, randoms=rnd)
    H, W = si.size.y, si.size.x
    bm = np.zeros((H, W, 3), np.uint8); ids = np.zeros((H, W, 4), np.int32)
    e.register_host(bm); e.register_host(ids)
    n0 = e.frames_streamed()
    e.render(si, eye, target, angles); e.readback(si, bitmap=bm, ids=ids)
    ref_bm, ref_ids = bm.copy(), ids.copy()
    e.render(si, (eye[0] + 300.0, eye[1], eye[2]), target, angles)              # streamed
    e.render(si, eye, target, angles)                                            # nobody read the previous one: not streamed
    assert e.frames_streamed() - n0 == 1
    e.readback(si, bitmap=bm, ids=ids)
    assert np.array_equal(bm, ref_bm) and np.array_equal(ids, ref_ids)
    e.unregister_host(bm); e.unregister_host(ids)
    e.render(si, (eye[0] + 300.0, eye[1], eye[2]), target, angles)
    e.synchronize()
    assert np.array_equal(bm, ref_bm)   # no kernel writes a buffer that is no longer registered
    bm2, ids2 = e.readback(si)
    assert not np.array_equal(bm2, ref_bm)
    e.close()


def test_host_container_streams_its_own_buffers():
    """SceneHost pins the frame and id buffers it owns (GPUKernel.cpp:344-360 allocates them once): with the reference's
    protocol (ids every frame) its second frame onwards is streamed, and equals the first."""
    sc, si, eye, target, angles, rnd, _ = gs.case_setup("mixed_full")
    si.maxPathTracingIterations = 1 << 30
    h = host.SceneHost(si)
    sc.replay(h)
    h.set_randoms(rnd, 0)
    h.set_camera(eye, target, angles)
    h.init_buffers()
    h.set_lazy_ids(False)
    e = engine.Engine.__new__(engine.Engine); e.lib = engine.load()
    n0 = e.frames_streamed()
    h.render_begin(0.0); h.render_end()
    bm1, ids1 = h.bitmap().copy(), h.primitive_ids().copy()
    for _ in range(3):
        h.render_begin(0.0); h.render_end()
        assert np.array_equal(bm1, h.bitmap()) and np.array_equal(ids1, h.primitive_ids())
    assert e.frames_streamed() - n0 == 3
    h.close()
