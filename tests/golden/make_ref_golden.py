"""Generates tests/golden/ref_vectors.npz by running the seeded cases of tests/refcases.py through
oracle/_ref/libmvsref.so — the REFERENCE'S OWN sources (Solver/SRTSolver.cpp, Camera/Camera.cpp,
SetUtils/*.cpp, Depth2Model/Depth2Model.cpp, Common/Utils.h) compiled unmodified from
/root/reference (`make -C oracle ref`).  These are outputs of the reference itself run here; they
travel to the GPU box, where /root/reference does not exist, and pin both the oracle restatement
(tests/test_ref_parity.py) and the CUDA path (tests/test_gpu_ref_parity.py).
Run (in the container that has /root/reference):  python tests/golden/make_ref_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from oracle import ref  # noqa: E402
from tests import refcases  # noqa: E402


class RefBackend:
    """oracle.ref with the call shapes refcases.run expects"""
    srt_estimate = staticmethod(ref.srt_estimate)
    srt_residual = staticmethod(ref.srt_residual)
    srt_triples = staticmethod(ref.srt_triples)
    project = staticmethod(ref.project)
    unproject = staticmethod(ref.unproject)
    pointset = staticmethod(ref.pointset)
    connected_keep = staticmethod(ref.connected_keep)
    depth2model = staticmethod(ref.depth2model)

    @staticmethod
    def srt_ransac(p, q, c1, c2, iters, seed, array, triples):
        return ref.srt_ransac(p, q, c1, c2, iters, seed, array)  # draws `triples` itself: srand(seed)


def main():
    ref.build()
    out = refcases.run(RefBackend, oracle.make_camera)
    path = Path(__file__).resolve().parent / "ref_vectors.npz"
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size, "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
