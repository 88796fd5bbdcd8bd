"""CUDA path (through the C ABI) against tests/golden/ref_vectors.npz — outputs of the REFERENCE'S
OWN code (oracle/_ref: Solver/SRTSolver.cpp, Camera/Camera.cpp, SetUtils/*.cpp,
Depth2Model/Depth2Model.cpp compiled unmodified; generator tests/golden/make_ref_golden.py) on the
seeded inputs of tests/refcases.py.  Rows S1-S6, A1, A7, §8(f)3 of SURVEY.md §8; further down the
same for the reference's Alignment / PartRecognition / PlyObj (ref_align_vectors.npz: rows A2-A7, f4)
and Deformation (ref_deform_vectors.npz: rows D2-D5, D4')."""
from pathlib import Path

import numpy as np
import pytest

from multiviewstitch_b200 import api
from tests import refcases

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden" / "ref_vectors.npz"


class CudaBackend:
    def __init__(self, ctx, golden):
        self.ctx, self.g = ctx, golden

    def _solver(self, p, q, c1, c2, iters=10):
        s = api.SRTSolver(iters, self.ctx)
        s.SetPrintFlag(False)
        s.SetInput(p, q, c1, c2)
        return s

    def srt_estimate(self, p, q, c1, c2):
        s = self._solver(p, q, c1, c2)
        out = s.EstimateTransform()
        s.close()
        return out

    def srt_residual(self, p, q, c1, c2, scale, R, t):
        s = self._solver(p, q, c1, c2)
        e = s.ResidualError(scale, R, t)
        s.close()
        return e

    def srt_triples(self, n, iters, seed):
        for name in ("noisy", "outliers"):
            tr = self.g[f"ransac_{name}_triples"]
            if len(tr) == iters and tr.max() < n:
                return tr
        raise KeyError((n, iters, seed))

    def srt_ransac(self, p, q, c1, c2, iters, seed, array, triples):
        s = self._solver(p, q, c1, c2, iters)
        out = s.EstimateTransformArray(triples) if array else s.EstimateTransformRansac(triples)
        s.close()
        return out

    def project(self, cam, pts):  # no stand-alone projection entry: covered through ResidualError
        return self.g["project_uv1"] if cam.t[0] == 0 and cam.t[2] == 0 else self.g["project_uv2"]

    def unproject(self, cam, uv, depth):
        _, dsp, _ = refcases.unproject_inputs()
        img = np.zeros((cam.h, cam.w), np.float32)
        img[uv[:, 1], uv[:, 0]] = dsp
        pts, valid = api.unproject_pixels(img, cam, uv, refcases.MIN_DSP, refcases.MAX_DSP, self.ctx)
        assert valid.all()
        return pts

    def pointset(self, pts):
        ps = api.PointSetUtils(self.ctx)
        ps.SetInput(pts)
        mn, mx = ps.GetBoundingBox()
        return ps.GetBarycenter(), (mn + mx) * 0.5, ps.CalcPivots(3)

    def connected_keep(self, n, facets):
        ids = np.zeros((n, 3))
        ids[:, 0] = np.arange(n)
        kept, _, _ = api.Alignment(self.ctx).RetainConnectRegion(ids, None, facets)
        keep = np.zeros(n, bool)
        keep[kept[:, 0].astype(np.int64)] = True
        return keep, -1

    def depth2model(self, d, cam, mn, mx, smooth):
        return api.depth2model(d, cam, mn, mx, smooth, self.ctx)


def test_cuda_matches_reference_vectors(ctx):
    """north_star bars: SRT parameters within 1e-6 relative (measured: ~1e-12), integer outputs
    (pixel-derived residuals, kept sets, texIndex, facets) exact."""
    want = dict(np.load(G))
    got = refcases.run(CudaBackend(ctx, want), api.make_camera)
    worst = 0.0
    for k, w in want.items():
        g = got[k]
        if k.endswith("_ncomp") or k.startswith("project_"):
            continue
        if w.dtype.kind in "iub":
            assert g.shape == w.shape and np.array_equal(g, w), k
        elif k.endswith("_pivots"):
            for c in range(3):
                s = np.sign(np.dot(g[:, c], w[:, c]))
                assert np.abs(s * g[:, c] - w[:, c]).max() <= 1e-9, k
        elif k.endswith("_res") or k.endswith("_res_planted"):
            # mean of integer pixel distances: any difference is a rounding-boundary flip
            assert abs(float(g) - float(w)) <= 1e-9 * max(1.0, float(w)), (k, float(g), float(w))
        else:
            scale = max(1.0, float(np.abs(w).max()))
            err = float(np.abs(g - w).max()) / scale
            worst = max(worst, err)
            assert err <= 1e-6, f"{k}: {err:.3e} relative"
    print(f"worst relative difference to the reference over all float outputs: {worst:.3e}")
    assert worst <= 1e-8


def test_cuda_array_overload_equals_ransac(ctx):
    """S6: EstimateTransform(double&, double[3][3], double[3]) (SRTSolver.cpp:256-270) is the RANSAC
    path with array outputs — identical numbers through mvs_srt_estimate_array, with given samples
    and with library-drawn ones."""
    want = dict(np.load(G))
    c1, c2 = refcases.cams(api.make_camera)
    p, q = refcases.srt_inputs()["outliers"]
    s = api.SRTSolver(200, ctx)
    s.SetPrintFlag(False)
    s.SetInput(p, q, c1, c2)
    tr = want["ransac_outliers_triples"]
    a = s.EstimateTransformArray(tr)
    b = s.EstimateTransformRansac(tr)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert np.abs(a[1] - want["ransac_outliers_array_R"]).max() <= 1e-9
    assert np.abs(a[2] - want["ransac_outliers_array_t"]).max() <= 1e-8
    s.SetIterationNum(64)
    a = s.EstimateTransformArray(None, seed=5)
    b = s.EstimateTransformRansac(api.srt_draw_triples(len(p), 64, 5))
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    s.close()


# ------------------------------------------------------------------ Alignment / PartRecognition / OBJ
GA = Path(__file__).resolve().parent / "golden" / "ref_align_vectors.npz"


def test_cuda_alignment_matches_reference_vectors(ctx):
    """Rows A2-A7 through the C ABI against tests/golden/ref_align_vectors.npz — outputs of the
    reference's OWN Alignment.cpp / PartRecognition.cpp (oracle/ref_shim_align.cpp).  Kept point sets,
    facets and labels exact; the transformed template within 1e-8 absolute (measured ~1e-13)."""
    want = dict(np.load(GA))
    from multiviewstitch_b200 import synth
    al = api.Alignment(ctx, api.align_params(dist_thres=refcases.DIST_THRES))
    # A7
    v1, f1 = synth.uv_sphere(10, 7)
    v2, f2 = synth.uv_sphere(14, 9)
    v = np.concatenate([v1, v2 + 5.0], 0)
    f = np.concatenate([f1, f2 + len(v1)], 0)
    p, n, ff = al.RetainConnectRegion(v, v.copy(), f)
    assert np.array_equal(p, want["rc_pts"]) and np.array_equal(n, want["rc_normals"])
    assert np.array_equal(ff, want["rc_facets"])
    # A4
    bv, bn, bf, nb = synth.body_on_ground()
    p, n, ff, gr = al.RemoveGround(bv, bn, bf)
    assert np.array_equal(p, want["rg_pts"]) and np.array_equal(n, want["rg_normals"])
    assert np.array_equal(ff, want["rg_facets"])
    assert np.abs(gr - want["rg_ground_ray"]).max() <= 1e-9
    # A5 on the reference's own initially aligned template
    sc = refcases.align_scene()
    body = sc["tv"][: -len(synth.grid_disk(0.7, 36, 0.0)[0])]
    T, s = want["ia_pose_T"], float(want["ia_pose_s"])
    s1 = s * sc["sv"] @ T[:3, :3].T + T[:3, 3]
    pr = api.PartRecognition(ctx)
    pr.SetInput(s1)
    pr.SetLabels(sc["sl"])
    assert np.array_equal(pr.PartRecog(body), want["pr_labels"])
    # Alignment::Align as a whole: RemoveGround, InitAlignment, PartRecog, neck offset, four limbs
    gs, gsn, gt, gtn, gtf, gtl, gg = al.Align(sc["sv"], sc["sn"], sc["sl"], sc["tv"], sc["tn"], sc["tf"],
                                              sc["view_ray"])
    assert np.array_equal(gt, want["al_tgt"]) and np.array_equal(gtn, want["al_tgt_normals"])
    assert np.array_equal(gtf, want["al_tgt_facets"])
    e_s, e_n = np.abs(gs - want["al_src"]).max(), np.abs(gsn - want["al_src_normals"]).max()
    print(f"Align vs the reference: template {e_s:.3e}, normals {e_n:.3e}")
    assert e_s <= 1e-8 and e_n <= 1e-8


def test_export_obj_bytes_equal_reference_writer(ctx, tmp_path):
    """§8(f)4: mvs_deform_export_obj against the reference's own Mesh::WriteObj (PlyObj.cpp:77-137)
    fed with the library's vertices, normals and facets: the files agree byte for byte (apart from
    the `# Object <name>` comment), and the reference's reader gets the mesh back."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not shipped")
    from multiviewstitch_b200 import synth
    v, f = synth.uv_sphere(16, 12, bumps=0.08)
    pts = v.astype(np.float32)
    d = api.Deformation(pts, synth.vertex_normals(v, f).astype(np.float32), f, ctx)
    api.exportOBJ(str(tmp_path / "gpu.obj"), d)
    ref.write_obj(tmp_path / "ref.obj", d.Vertices(), d.Normals(), d.Facets())
    a = [l for l in (tmp_path / "gpu.obj").read_text().splitlines() if not l.startswith("# Object")]
    b = [l for l in (tmp_path / "ref.obj").read_text().splitlines() if not l.startswith("# Object")]
    assert a == b
    p2, n2, f2 = ref.read_obj(tmp_path / "gpu.obj", len(pts), len(f))
    assert np.array_equal(f2, f) and np.abs(p2 - pts).max() < 1e-5


# ------------------------------------------------------------------ Deformation D2-D5, D4'
GD = Path(__file__).resolve().parent / "golden" / "ref_deform_vectors.npz"


def test_cuda_deformation_graph_matches_reference_vectors(ctx):
    """Rows D2-D5 and D4' through the C ABI against tests/golden/ref_deform_vectors.npz — outputs of
    the reference's OWN Deformation.cpp (oracle/ref_shim_deform.cpp; exact nearest neighbours in place
    of FLANN, CGAL's solver replaced by a recorder of the constraints it is handed): sampled node
    ids and node graph exact; control positions after the correspondence rule and the two smoothing
    sweeps within 1e-12; the Deform(threshold) constraints exact."""
    want = dict(np.load(GD))
    sc = refcases.deform_scene()
    d = api.Deformation(sc["pts"], sc["nrm"], sc["faces"], ctx)
    ids = d.UniformSampling()
    assert np.array_equal(ids, want["d2_ids"])
    assert np.array_equal(d.KNearestNeighbor(8), want["d3_graph"])
    d.SetTarget(sc["tv"], sc["tn"])
    worst = 0.0
    for k, (pl, pd) in enumerate(refcases.DEFORM_THRESHOLDS):
        c, v = d.NodeTargets(pl, pd, 2)
        worst = max(worst, float(np.abs(c - want[f"d4_controls_{k}"]).max()))
    print(f"node control positions vs the reference: max abs diff {worst:.3e}")
    assert worst <= 1e-12
    for k, thr in enumerate(refcases.DEFORM_D4P_THRESHOLDS):
        c, v = d.NodeTargetsThreshold(thr)
        assert np.array_equal(ids[v > 0], want[f"d4p_vertices_{k}"])
        assert np.array_equal(c[v > 0], want[f"d4p_targets_{k}"])


# ------------------------------------------------------------------ Processor::CheckConsistencyCore
def test_cuda_check_consistency_matches_reference_vectors(ctx):
    """§8(f)1: mvs_check_consistency against the depth map the reference's own
    Processor::CheckConsistencyCore (Processor.cpp:66-115) leaves — identical."""
    import oracle
    want = dict(np.load(Path(__file__).resolve().parent / "golden" / "ref_processor_vectors.npz"))

    def m2d(verts, faces, poses, K, w, h):
        return oracle.model2depth(verts, faces, [oracle.make_camera(K, R, t, w, h) for R, t in poses])

    cs, maps = refcases.consistency_inputs(api.make_camera, m2d)
    g = api.check_consistency(maps[1], cs[1], maps[[0, 2]], [cs[0], cs[2]], refcases.MIN_DSP, 0.5, 4.0, ctx)
    assert np.array_equal(np.asarray(g, np.float32), want["cc_depth"])


def test_cuda_remove_outliers_matches_reference_vectors(ctx):
    """§8(f)1 directly: mvs_remove_outliers_samples, handed the hypothesis samples the reference drew
    from rand() round by round (stored next to its results), keeps exactly the matches the reference's
    own Processor::RemoveOutliers keeps and reports its mean pixel error."""
    want = dict(np.load(Path(__file__).resolve().parent / "golden" / "ref_processor_vectors.npz"))
    p, q, _ = refcases.outlier_inputs()
    c1, c2 = refcases.cams(api.make_camera)
    for seed, pe, ar in refcases.REMOVE_OUTLIER_CASES:
        stored = [want[k] for k in sorted(want) if k.startswith(f"ro_{seed}_round")]
        served = []

        def draw(size, iters):
            tr = stored[len(served)]
            assert tr.shape == (iters, 3) and tr.max() < size
            served.append(tr)
            return tr

        keep, err, _ = api.remove_outliers_samples(p, q, c1, c2, draw, 3, 200, pe, ar, ctx)
        assert len(served) == len(stored)
        assert np.array_equal(keep, want[f"ro_{seed}_keep"])
        assert abs(err - float(want[f"ro_{seed}_err"])) <= 1e-9 * max(1.0, float(want[f"ro_{seed}_err"]))
