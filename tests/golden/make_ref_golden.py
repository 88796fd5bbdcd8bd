"""Generates tests/golden/ref_vectors.npz by running the seeded cases of tests/refcases.py through
oracle/_ref/libmvsref.so — the REFERENCE'S OWN sources (Solver/SRTSolver.cpp, Camera/Camera.cpp,
SetUtils/*.cpp, Depth2Model/Depth2Model.cpp, Common/Utils.h) compiled unmodified from
/root/reference (`make -C oracle ref`) — and tests/golden/ref_align_vectors.npz from the reference's
Alignment.cpp, PartRecognition.cpp and PlyObj.cpp (oracle/ref_shim_align.cpp), and
tests/golden/ref_deform_vectors.npz from its Deformation.cpp (oracle/ref_shim_deform.cpp), and
tests/golden/ref_processor_vectors.npz from Processor::RemoveOutliers / CheckConsistencyCore
(oracle/ref_shim_processor.cpp).  These are outputs of the reference itself run here; they
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


class RefAlignBackend:
    """oracle.ref's Alignment / PartRecognition / PlyObj entry points with the call shapes
    refcases.run_align expects.  The reference works in the current directory (it reads
    ./Template/part/parts and writes debug meshes): a scratch directory per call."""

    def __init__(self, workdir):
        self.wd = Path(workdir)
        self.wd.mkdir(parents=True, exist_ok=True)

    retain_connected = staticmethod(ref.retain_connected)
    init_alignment = staticmethod(ref.init_alignment)
    part_recog = staticmethod(ref.part_recog)
    vertex_normals = staticmethod(ref.mesh_vertex_normals)

    def remove_ground(self, p, n, f, dist_thres):
        return ref.remove_ground(p, n, f, dist_thres, self.wd)

    @staticmethod
    def local_core(src, sl, tgt, tl, sel, label):
        # the reference is handed the limb's points only (Alignment.cpp:326-371)
        sl, tl = np.asarray(sl), np.asarray(tl)
        ms, mt = np.isin(sl, sel), np.isin(tl, sel)
        return ref.local_alignment_core(src[ms], sl[ms], tgt[mt], tl[mt], label, label)

    def align(self, sv, sn, sf, sl, tv, tn, tf, view_ray, dist_thres):
        return ref.align(sv, sn, sf, sl, tv, tn, tf, view_ray, dist_thres, self.wd)


class RefDeformBackend:
    """oracle.ref's Deformation entry points with the call shapes refcases.run_deform expects"""

    def __init__(self, workdir):
        self.wd = Path(workdir)
        self.wd.mkdir(parents=True, exist_ok=True)

    @staticmethod
    def _m(sc):
        return sc["pts"].astype(np.float64), sc["nrm"].astype(np.float64), sc["faces"]

    def uniform_sampling(self, sc):
        return ref.deform_uniform_sampling(*self._m(sc), self.wd)

    def node_graph(self, sc, ids, K):
        return ref.deform_node_graph(*self._m(sc), ids, K)[0]

    def node_targets(self, sc, ids, graph, pl, pd):
        v, t = ref.deform_node_targets(*self._m(sc), ids, sc["tv"].astype(np.float64),
                                       sc["tn"].astype(np.float64), pl, pd, self.wd)
        assert np.array_equal(v, ids)  # one constraint per node, in node order (:383-391)
        return t

    def node_targets_threshold(self, sc, ids, thr):
        return ref.deform_node_targets_threshold(*self._m(sc), ids, sc["tv"].astype(np.float64),
                                                 sc["tn"].astype(np.float64), thr)

    @staticmethod
    def mesh_is_valid(pts, faces):
        return int(ref.deform_mesh_is_valid(pts.astype(np.float64), faces))


def oracle_model2depth(verts, faces, poses, K, w, h):
    """the synthetic depth maps of the consistency case (inputs, not reference outputs)"""
    return oracle.model2depth(verts, faces, [oracle.make_camera(K, R, t, w, h) for R, t in poses])


class RefProcessorBackend:
    """oracle.ref's Processor entry points with the call shapes refcases.run_processor expects"""

    @staticmethod
    def remove_outliers(p, q, c1, c2, pixel_err, adapt_ratio, seed):
        keep, err, _ = ref.remove_outliers(p, q, c1, c2, pixel_err, adapt_ratio, seed)
        # the samples the reference drew, round by round: replay its rand() stream through the round
        # loop (the number of matches left after each round decides how the next ones are drawn)
        import ctypes as C
        rounds = []

        @C.CFUNCTYPE(None, C.c_int64, C.c_int, C.POINTER(C.c_int32), C.c_void_p)
        def draw(size, iters, out, user):
            ref.lib().ref_srt_triples_next(C.c_int64(size), C.c_int(iters), out, None)
            rounds.append(np.ctypeslib.as_array(out, shape=(iters, 3)).copy())

        ref.lib().ref_srand(C.c_uint(seed))
        k2, e2, _ = oracle.remove_outliers_cb(p, q, c1, c2, 3, 200, pixel_err, adapt_ratio, draw)
        assert np.array_equal(k2, keep) and e2 == err, "oracle round loop left the reference's path"
        return keep, err, rounds

    @staticmethod
    def check_consistency(depth, cam, ref_depths, ref_cams, min_dsp, max_dsp, reproj_err):
        return ref.check_consistency(np.asarray(depth, np.float64), cam, np.asarray(ref_depths, np.float64),
                                     ref_cams, min_dsp, max_dsp, reproj_err)


def main():
    import tempfile
    ref.build()
    out = refcases.run(RefBackend, oracle.make_camera)
    path = Path(__file__).resolve().parent / "ref_vectors.npz"
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size, "bytes,", len(out), "arrays")
    with tempfile.TemporaryDirectory() as wd:
        out = refcases.run_align(RefAlignBackend(wd))
    path = Path(__file__).resolve().parent / "ref_align_vectors.npz"
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size, "bytes,", len(out), "arrays")
    with tempfile.TemporaryDirectory() as wd:
        out = refcases.run_deform(RefDeformBackend(wd))
    path = Path(__file__).resolve().parent / "ref_deform_vectors.npz"
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size, "bytes,", len(out), "arrays")
    out = refcases.run_processor(RefProcessorBackend, oracle.make_camera, oracle_model2depth)
    path = Path(__file__).resolve().parent / "ref_processor_vectors.npz"
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size, "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
