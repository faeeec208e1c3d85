"""Generates tests/golden/*.npz by running the UNMODIFIED REFERENCE (oracle/_ref/libsolr_ref_cpu.so: the
reference's host library + its CUDA engine source compiled for the host, see oracle/ref_build/) on the
cases of tests/golden_scenes.py.  Run in the build container (needs /root/reference to have been built:
`make -C oracle ref`); the .npz files travel, the reference does not.

Stored per case: the flattened scene arrays' SHA-256 (what compactBoxes produced), and after the last
frame the id buffer, the RGB8 bitmap and the float accumulation buffer, bit for bit."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import golden_scenes as gs  # noqa: E402
import refh  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    only = sys.argv[1:]
    for name in gs.CASES:
        if only and not any(name.startswith(o) for o in only):
            continue
        sc, si, eye, target, angles, rnd, frames = gs.case_setup(name)
        pp = gs.case_post(name)
        r = refh.RefScene(si, "cpu")
        sc.replay(r)
        a = r.arrays()
        for it in frames:
            si.pathTracingIteration = it
            bm, ids, post = r.render(si, eye, target, angles, randoms=rnd, post_info=pp, block=(16, 8))
        li = a["lightInformation"].reshape(-1, 48)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), ids=ids, bitmap=bm, post=post,
                            boxes_sha=sha(a["boxes"]), primitives_sha=sha(a["primitives"]),
                            materials_sha=sha(a["materials"]), lights_sha=sha(np.concatenate([li[:, :20], li[:, 32:]], 1)),
                            nbBoxes=a["nbBoxes"], nbPrimitives=a["nbPrimitives"], frames=np.array(frames))
        print(name, "boxes", a["nbBoxes"], "hit px", int((ids[..., 0] >= 0).sum()), "mean rgb", float(bm.mean()))
        r.close()


if __name__ == "__main__":
    main()
