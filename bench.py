#!/usr/bin/env python
"""bench.py — headline benchmark of the non-rigid alignment hot path (BASELINE.json metric).

Workload (BASELINE.json configs[2], "C3"): one closed 1,003,688-vertex cube-sphere mesh with bumps
(near-uniform density), a 4096-node deformation graph (farthest-point sampling, k = 4 node weights,
8-neighbour node graph), target "scan" = the same shape re-tessellated (turned, 2.7 % finer) and warped
by a smooth bend+twist of 5 % of the bbox diagonal, exact-NN correspondences re-derived in every
Gauss-Newton iteration, 10 GN iterations x 20 block-Jacobi PCG iterations, then the vertex update
(deformed vertices + recomputed normals).

A "step" is one such 10-iteration solve starting from the identity node state.
  value  = GN iterations / s with every input resident in HBM: K steps, each ONE library call
           (mvs_deform_solve), device-timed, per-family profiling OFF; `families` and `roofline` come
           from a second, separately profiled pass over the same K steps
  e2e    = the same metric through the host-buffer C ABI a reference caller would use:
           mvs_deform_set_target_f32(host) -> mvs_deform_solve -> mvs_deform_get_vertices_f32(host),
           H2D/D2H inside the timed region
  --gpus N: one process per GPU.  The headline line is N independent C3 solves, every rank on its own
           copy of the same mesh (the reference's independent objects, no data-path collective,
           scaling "weak"); at EVERY N the line also
           carries north_star's two partitions: secondary.c4_pairs (configs[3]: 28 view pairs sharded
           pair -> rank, one mvs_stitch_pairs call per rank) and secondary.c5_shard (configs[4]: ONE
           5M-vertex / 20K-node graph vertex-sharded over the ranks with an NCCL all-reduce of the
           fixed-point J^T J / J^T r blocks, checked bitwise against the unsharded solve).
  --mode shard / --mode pairs: those two partitions as the headline line of their own run.
  --impl reference: the CPU oracle (oracle/, single thread like the reference) on the same
           workload; a step = one full solve, and `steps` reports the solves actually timed
           inside --ref-budget-s.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "deformation_gn_iters_per_s"
UNIT = "GN iters/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="objects", choices=["objects", "shard", "pairs"])
    ap.add_argument("--pairs", type=int, default=28, help="view pairs of --mode pairs (C(8,2) = 28)")
    ap.add_argument("--pair-workers", type=int, default=0,
                    help="--mode pairs: host threads per rank, each with its own context/stream, "
                         "working through the rank's pairs concurrently (independent units); "
                         "0 = min(4, host cores / ranks)")
    ap.add_argument("--side", type=int, default=409, help="cube-sphere grid side (verts = 6*side^2+2)")
    ap.add_argument("--nodes", type=int, default=4096)
    ap.add_argument("--gn-iters", type=int, default=10)
    ap.add_argument("--pcg-iters", type=int, default=20)
    ap.add_argument("--ref-gn-iters", type=int, default=0,
                    help="GN iterations of the in-line cpu_baseline solve (0 = --gn-iters)")
    ap.add_argument("--ref-budget-s", type=float, default=100.0)
    ap.add_argument("--c5-side", type=int, default=913, help="secondary.c5_shard mesh (5,001,416 verts)")
    ap.add_argument("--c5-nodes", type=int, default=20000)
    ap.add_argument("--c5-steps", type=int, default=3)
    ap.add_argument("--c5-cpu", action="store_true",
                    help="also run the CPU oracle on the C5 workload (about two minutes) and report "
                         "the distance of the GPU result to it")
    ap.add_argument("--no-c4", action="store_true")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the NN-corr/s + SRT closed-form lines (configs[1], 1e7 points)")
    ap.add_argument("--c2-points", type=int, default=10_000_000)
    return ap.parse_args()


def workload(args, seed_shift=0):
    """Template = cube-sphere mesh with low-frequency bumps (near-uniform vertex density, closed,
    genus 0); target "scan" = the same shape sampled with its own, turned and finer tessellation
    (a scan never shares the template's lattice) and warped by a smooth bend + twist of 5 % of the
    bounding-box diagonal."""
    from multiviewstitch_b200 import synth
    verts, faces = synth.cube_sphere(args.side)
    tv, _ = synth.cube_sphere(int(round(args.side * 1.027)), rot=synth.axis_angle([1, 2, 3], 0.7))
    if seed_shift:
        # independent objects per rank: a different similarity of the same shape
        R = synth.axis_angle([1, 2 + seed_shift, 3], 0.1 * seed_shift)
        verts = (1.0 + 0.01 * seed_shift) * verts @ R.T
        tv = (1.0 + 0.01 * seed_shift) * tv @ R.T
    pts = verts.astype(np.float32)
    nrm = synth.vertex_normals(verts, faces).astype(np.float32)
    tgt = synth.bend_twist(tv, 0.05).astype(np.float32)
    return pts, nrm, faces, tgt


def workload_name(args, n):
    return (f"C3: {n}-vertex mesh, {args.nodes}-node graph (k=4, K=8), NN correspondences per "
            f"iteration, {args.gn_iters} GN x {args.pcg_iters} PCG + vertex update")


# ------------------------------------------------------------------------- clocks sampler
class Clocks:
    """SM clock, power and throttle reasons of one GPU while the timed region runs.  Polls NVML (the
    library behind nvidia-smi) every 5 ms from a thread started BEFORE the warm-up, so the samples
    reported are the ones whose time stamp lies inside the timed region (`window`); falls back to an
    `nvidia-smi -lms` loop when the NVML binding is not importable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason* bits (nvml.h)
    MASKS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20),
             ("hw_thermal_slowdown", 0x40))

    def __init__(self, index):
        self.index = index
        self.samples = []  # (time, sm MHz, max sm MHz, power W or None, set of reasons)
        self.proc = None
        self.source = None
        self.t0 = self.t1 = None
        self._halt = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self._nv = pynvml
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self._nv, self._h
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        while not self._halt.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = None
                self.samples.append((time.perf_counter(), sm, mx, pw,
                                     {n for n, m in self.MASKS if bits & m}))
            except Exception:
                pass
            self._halt.wait(0.005)

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                self.samples.append((time.perf_counter(), float(f[1]), float(f[2]), float(f[3]),
                                     {n for k, n in enumerate(names) if f[4 + k].lower().startswith("active")}))
            except ValueError:
                continue

    def window(self, t0, t1):
        """the timed region, in time.perf_counter() seconds"""
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.source == "nvml":
            self._halt.set()
            self.thread.join(timeout=1)
        else:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = list(self.samples)
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        use = inside or rows
        sm = [r[1] for r in use]
        mx = [r[2] for r in use if r[2] is not None]
        pw = [r[3] for r in use if r[3] is not None]
        reasons = set()
        for r in use:
            reasons |= r[4]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "in_timed_region": len(inside), "source": self.source,
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------- reference (CPU) arm
def cpu_solve(args, pts, tgt, gn_iters):
    import oracle
    t0 = time.perf_counter()
    ids = oracle.sample_fps(pts, args.nodes, 0)
    nodes = pts[ids]
    widx, w = oracle.node_weights(pts, nodes, 4)
    graph = oracle.node_graph(nodes, 8)
    t_setup = time.perf_counter() - t0
    prm = oracle.gn_params(gn_iters=gn_iters, pcg_iters=args.pcg_iters, corr_mode=1)
    t0 = time.perf_counter()
    x, out, e = oracle.gn_solve(pts, None, ids, widx, w, graph, prm, tpts=tgt)
    t_solve = time.perf_counter() - t0
    return t_setup, t_solve, out, (ids, widx, w, graph)


def cpu_solve_openmp(args, pts, tgt, gn_iters, graph_parts):
    """Courtesy number, NOT the reference: the same oracle solve compiled with -fopenmp on every host
    core (the reference and the parity oracle are single-threaded).  None when libgomp is absent."""
    import oracle
    if oracle.lib_mt() is None:
        return None
    ids, widx, w, graph = graph_parts
    prm = oracle.gn_params(gn_iters=gn_iters, pcg_iters=args.pcg_iters, corr_mode=1)
    t0 = time.perf_counter()
    oracle.gn_solve(pts, None, ids, widx, w, graph, prm, tpts=tgt, mt=True)
    t_solve = time.perf_counter() - t0
    return {"value": gn_iters / t_solve, "unit": UNIT, "cores": os.cpu_count(), "kind": "port-openmp",
            "note": "courtesy multi-thread build of the oracle port (queries, assembly and sparse "
                    "products in parallel, kd-tree build serial); the reference is single-threaded"}


def base_config(args, n, nt):
    """The `config` object — identical on both arms (the reference arm runs the same workload)."""
    return {"workload": workload_name(args, n), "verts": int(n), "target_points": int(nt),
            "nodes": int(args.nodes), "gn_iters": int(args.gn_iters), "pcg_iters": int(args.pcg_iters),
            "mode": args.mode,
            "step": f"one {args.gn_iters}-iteration solve from the identity node state + vertex update",
            "l2": "per-iteration working set ~100 MB (> L2 with the target); the GPU arm also writes a "
                  "512 MiB flush buffer between timed steps"}


def run_reference(args):
    """The reference's CPU path (the oracle port: the reference's Deformation stage needs CGAL and
    FLANN and cannot be compiled here) on the SAME workload with all the host threads it can use —
    one, like the reference.  A step = one full 10-iteration solve (about 7 s); the run is bounded
    by --ref-budget-s and `steps` reports the solves actually timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pts, nrm, faces, tgt = workload(args)
    import oracle
    ids = oracle.sample_fps(pts, args.nodes, 0)
    nodes = pts[ids]
    widx, w = oracle.node_weights(pts, nodes, 4)
    graph = oracle.node_graph(nodes, 8)
    prm = oracle.gn_params(gn_iters=args.gn_iters, pcg_iters=args.pcg_iters, corr_mode=1)
    times = []
    t_begin = time.perf_counter()
    warm = 1 if args.warmup > 0 else 0  # one pass pages the arrays in; the CPU has no JIT to warm
    for s in range(warm + args.steps):
        if times and time.perf_counter() - t_begin > args.ref_budget_s:
            break  # bounded sample: the whole run stays within a few minutes
        t0 = time.perf_counter()
        x, out, _ = oracle.gn_solve(pts, None, ids, widx, w, graph, prm, tpts=tgt)  # incl. vertex update
        dt = time.perf_counter() - t0
        if s >= warm:
            times.append(dt)
    tot = sum(times)
    val = args.gn_iters * len(times) / tot
    sample = (f"{len(times)} full solve(s) of {args.gn_iters} GN x {args.pcg_iters} PCG iterations on the "
              f"whole {len(pts)}-vertex / {args.nodes}-node workload (kd-tree built once per solve, exact "
              f"NN per iteration, deformed vertices evaluated at the end), 1 warm-up solve, "
              f"{args.steps} requested, bounded by {args.ref_budget_s:.0f} s")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "steps_requested": args.steps, "warmup": warm,
            "ms_per_step": 1e3 * tot / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": base_config(args, len(pts), len(tgt)),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": sample,
                             "note": "single thread because the reference is single-threaded; the "
                                     "reference's Deformation stage (CGAL + FLANN, MSVC) cannot be built "
                                     "here, so this leg is the oracle port; oracle/_ref (the reference's "
                                     "own SRT/Camera/PointSet/Depth2Model sources) pins the port"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cores": os.cpu_count(),
            "wall_s": time.perf_counter() - t_begin}
    try:  # courtesy number on every host core; `value` stays the single-thread reference arm
        line["cpu_baseline"]["openmp_courtesy"] = cpu_solve_openmp(
            args, pts, tgt, args.gn_iters, (ids, widx, w, graph))
    except Exception as exc:
        line["cpu_baseline"]["openmp_courtesy"] = {"error": repr(exc)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------- B200 arm
ALGO_BYTES_DOC = {
    # per GN iteration, n vertices, nt target points, m nodes, nnz pattern blocks
    "gn_deform_corr": "n*(16 v + 16 idx + 16 w + 16 resid) + nt*16 target cells",
    "gn_assemble_fit": "n*(16 v + 16 w + 16 resid) + runs*(64 node positions + 16 tile map) + "
                       "((nnz+m)/2*16 + m*12)*8 upper-triangle fixed-point system",
    "pcg_solve": "iters*((nnz*64 + m*324) matrix + m*576 Minv + 10*m*48 vectors)",
}


NRUNS = 0  # tuple runs of the benchmarked graph (set once the graph is built)


def algo_bytes(name, n, nt, m, nnz, pcg_iters):
    if name == "gn_deform_corr":
        return n * 64 + nt * 16
    if name == "gn_assemble_fit":
        return n * 48 + NRUNS * 80 + ((nnz + m) // 2 * 16 + m * 12) * 8
    if name == "pcg_solve":
        return pcg_iters * (nnz * 64 + m * 324 + m * 576 + 10 * m * 48)
    if name == "deform_final":
        return n * (48 + 16) + 0
    return None


def secondary_c2(args, api, ctx, stream, peak):
    """configs[1] (C2): exact NN correspondences and the SRT closed form at n = 1e7 float4 points,
    device resident: NN-corr/s (the second half of BASELINE.json's metric) and both kernels'
    fraction of the HBM roofline from their algorithmic bytes (40 B/correspondence, 64 B/pair)."""
    import torch
    from multiviewstitch_b200 import synth
    n = args.c2_points
    g = torch.Generator(device="cuda").manual_seed(1234)
    p = torch.randn((n, 3), device="cuda", generator=g)
    p = p / p.norm(dim=1, keepdim=True) + 1e-3 * torch.randn((n, 3), device="cuda", generator=g)
    p[:, 2] += 5.0
    R = torch.tensor(synth.PLANTED_R, dtype=torch.float32, device="cuda")
    t = torch.tensor(synth.PLANTED_T, dtype=torch.float32, device="cuda")
    q = synth.PLANTED_S * p @ R.T + t + 1e-3 * torch.randn((n, 3), device="cuda", generator=g)
    p4 = torch.zeros((n, 4), device="cuda"); p4[:, :3] = p
    q4 = torch.zeros((n, 4), device="cuda"); q4[:, :3] = q
    shuf = torch.randperm(n, device="cuda", generator=g)
    q4s = q4[shuf].contiguous()
    query = torch.zeros((n, 4), device="cuda")
    query[:, :3] = synth.PLANTED_S * p @ R.T + t            # p' = s R p + t against shuffled q
    idx_out = torch.empty(n, dtype=torch.int32, device="cuda")
    d2_out = torch.empty(n, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    out = {}
    with torch.cuda.stream(stream):
        index = api.NNIndex(ctx)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        index.build_dev(q4s.data_ptr(), n)
        index.knn_dev(query.data_ptr(), n, 1, idx_out.data_ptr(), d2_out.data_ptr())  # warm-up (+ pyramid)
        best = None
        for rep in range(3):
            ev[0].record(stream)
            index.knn_dev(query.data_ptr(), n, 1, idx_out.data_ptr(), d2_out.data_ptr())
            ev[1].record(stream)
            stream.synchronize()
            ms = ev[0].elapsed_time(ev[1])
            best = ms if best is None else min(best, ms)
        # size-independent checks: the reported d2 is the distance to the reported index, and no
        # farther than the planted partner
        inv = torch.empty_like(shuf); inv[shuf] = torch.arange(n, device="cuda")
        got = q4s[idx_out.long(), :3] - query[:, :3]
        d_got = (got[:, 0] * got[:, 0] + got[:, 1] * got[:, 1]) + got[:, 2] * got[:, 2]
        par = q4s[inv, :3] - query[:, :3]
        d_par = (par[:, 0] * par[:, 0] + par[:, 1] * par[:, 1]) + par[:, 2] * par[:, 2]
        ok = bool(torch.equal(d_got, d2_out)) and bool((d2_out <= d_par).all())
        # exactness spot check at full size: brute force over ALL n targets for 4096 random queries
        # (same FMA-free float32 distance, lowest index on ties)
        sel = torch.randint(0, n, (4096,), device="cuda", generator=g)
        bf_idx = torch.empty(4096, dtype=torch.int64, device="cuda")
        bf_d2 = torch.empty(4096, dtype=torch.float32, device="cuda")
        tx, ty, tz = q4s[:, 0].contiguous(), q4s[:, 1].contiguous(), q4s[:, 2].contiguous()
        for c0 in range(0, 4096, 64):
            qs = query[sel[c0:c0 + 64], :3]
            dx = tx[None, :] - qs[:, 0:1]
            dy = ty[None, :] - qs[:, 1:2]
            dz = tz[None, :] - qs[:, 2:3]
            dd = (dx * dx + dy * dy) + dz * dz
            mn = dd.min(dim=1, keepdim=True).values
            first = torch.where(dd == mn, torch.arange(n, device="cuda")[None, :],
                                torch.full((1, 1), n, device="cuda")).min(dim=1).values
            bf_idx[c0:c0 + 64] = first
            bf_d2[c0:c0 + 64] = mn[:, 0]
            del dx, dy, dz, dd
        brute_ok = bool(torch.equal(bf_idx, idx_out[sel].long())) and bool(torch.equal(bf_d2, d2_out[sel]))
        out["nn_corr"] = {"value": n / (best * 1e-3), "unit": "NN-corr/s", "points": n,
                          "ms": best, "binning_included": True,
                          "GBps": 40.0 * n / (best * 1e-3) / 1e9,
                          "frac_hbm": 40.0 * n / (best * 1e-3) / 1e9 / peak,
                          "self_consistent": ok, "brute_force_4096_queries_exact": brute_ok,
                          "own_partner_fraction": float((idx_out.long() == inv).float().mean())}
        srt = api.SRTSolver(10, ctx)
        srt.SetPrintFlag(False)
        srt.SetInputDev(p4.data_ptr(), q4.data_ptr(), n)
        srt.EstimateTransform()
        best = None
        for rep in range(3):
            ev[2].record(stream)
            sc, Rr, tt = srt.EstimateTransform()
            ev[3].record(stream)
            stream.synchronize()
            ms = ev[2].elapsed_time(ev[3])
            best = ms if best is None else min(best, ms)
        err = max(abs(sc - synth.PLANTED_S), float(np.abs(Rr - synth.PLANTED_R).max()),
                  float(np.abs(tt - synth.PLANTED_T).max()))
        out["srt_closed_form"] = {"value": n / (best * 1e-3), "unit": "pairs/s", "pairs": n, "ms": best,
                                  "GBps": 64.0 * n / (best * 1e-3) / 1e9,
                                  "frac_hbm": 64.0 * n / (best * 1e-3) / 1e9 / peak,
                                  "max_abs_error_vs_planted": err,
                                  "note": "includes the D2H read of (s,R,t); noise 1e-3 on 1e7 pairs"}
        index.close(); srt.close()
        # configs[1] is a sweep: the same two measurements on the first n points of the set
        sweep = []
        for ns in (10_000, 100_000, 1_000_000):
            if ns >= n:
                continue
            tq = q4s[:ns].contiguous()
            qq = query[:ns].contiguous()
            ix = api.NNIndex(ctx)
            ix.build_dev(tq.data_ptr(), ns)
            ix.knn_dev(qq.data_ptr(), ns, 1, idx_out.data_ptr(), d2_out.data_ptr())
            best_nn = None
            for rep in range(3):
                ev[0].record(stream)
                ix.knn_dev(qq.data_ptr(), ns, 1, idx_out.data_ptr(), d2_out.data_ptr())
                ev[1].record(stream)
                stream.synchronize()
                ms = ev[0].elapsed_time(ev[1])
                best_nn = ms if best_nn is None else min(best_nn, ms)
            ix.close()
            sr = api.SRTSolver(10, ctx)
            sr.SetPrintFlag(False)
            sr.SetInputDev(p4.data_ptr(), q4.data_ptr(), ns)
            sr.EstimateTransform()
            best_srt = None
            for rep in range(3):
                ev[2].record(stream)
                sr.EstimateTransform()
                ev[3].record(stream)
                stream.synchronize()
                ms = ev[2].elapsed_time(ev[3])
                best_srt = ms if best_srt is None else min(best_srt, ms)
            sr.close()
            sweep.append({"points": ns, "nn_ms": best_nn, "nn_corr_per_s": ns / (best_nn * 1e-3),
                          "srt_ms": best_srt, "srt_pairs_per_s": ns / (best_srt * 1e-3)})
        sweep.append({"points": n, "nn_ms": out["nn_corr"]["ms"], "nn_corr_per_s": out["nn_corr"]["value"],
                      "srt_ms": out["srt_closed_form"]["ms"],
                      "srt_pairs_per_s": out["srt_closed_form"]["value"]})
        out["c2_sweep"] = sweep
    return out


def secondary_arap(args, api, ctx, stream, peak, d, pts, faces):
    """ARAP mode of Deform() (what the reference runs through CGAL) on the
    benchmarked mesh with its nodes as control vertices: time of a fixed 2 x 50 (local/global x CG)
    solve, the CG iteration's fraction of the HBM roofline, and the oracle on the same inputs."""
    import torch
    from multiviewstitch_b200 import synth
    ids = d.Nodes()
    R = synth.axis_angle([0, 0, 1], 0.2)
    ctrl = pts[ids].astype(np.float64) @ R.T * 1.03
    prm = api.arap_params(iters=2, cg_iters=50, tol=0.0, cg_tol=0.0)
    out = {}
    with torch.cuda.stream(stream):
        d.ARAP(ctrl, prm)  # warm-up: builds the cotangent CSR
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ctx.prof_enable(True)
        ctx.prof_reset()
        ev[0].record(stream)
        energy, it = d.ARAP(ctrl, prm)
        ev[1].record(stream)
        stream.synchronize()
        prof = ctx.prof()
        ctx.prof_enable(False)
        ms = ev[0].elapsed_time(ev[1])
        got = d.Vertices()
    n = len(pts)
    nnz = 6 * len(faces)  # directed edges of a closed manifold mesh = 3 per facet x 2
    cg_ms, cg_launches = prof.get("arap_cg_iter", (0.0, 0))
    iters_cg = max(cg_launches // 4, 1)
    cg_bytes = nnz * 8 + n * 204
    out = {"verts": n, "control_vertices": int(len(ids)), "local_global_iters": int(it),
           "cg_iters_per_global_step": 50, "ms": ms, "energy": energy,
           "cg_iteration_ms": cg_ms / iters_cg,
           "cg_iteration_GBps": cg_bytes / (cg_ms / iters_cg * 1e-3) / 1e9,
           "cg_iteration_frac_hbm": cg_bytes / (cg_ms / iters_cg * 1e-3) / 1e9 / peak,
           "cg_bytes_formula": "nnz*8 (col, w) + n*(40 spmv + 116 update + 48 direction)",
           "families_ms": {k: round(v[0], 4) for k, v in prof.items() if k.startswith(("arap", "vertex"))}}
    if not args.no_cpu_baseline:
        import oracle
        t0 = time.perf_counter()
        want, e_ref, _ = oracle.arap(pts, faces, ids, ctrl, iters=2, tol=0.0, cg_iters=50, cg_tol=0.0)
        dt = time.perf_counter() - t0
        diag = float(np.linalg.norm(pts.max(0) - pts.min(0)))
        out["cpu_oracle_s"] = dt
        out["cpu_cores"] = 1
        out["max_abs_diff_vs_oracle_over_bbox_diag"] = float(np.abs(got - want).max() / diag)
    return out


# ------------------------------------------------------------------------- C4: view pairs
def build_pair_work(api, pipeline, ctx, units):
    """Setup (untimed): the two depth maps + lattice matches of every pair this rank owns."""
    work = []
    m2d = api.Model2Depth(ctx)
    for u in units:
        sc = pipeline.make_scene(side=72, view_angle_deg=10.0 + 1.5 * u, seed=0)
        ca, cb = api.make_camera(*sc["cam_a"]), api.make_camera(*sc["cam_b"])
        m2d.SetInput([sc["verts_a"], sc["verts_b"]], [None, None], [sc["faces"], sc["faces"]], [[ca], [cb]])
        da, db = m2d.render(0)[0], m2d.render(1)[0]
        pa, va = api.unproject_depth(da, ca, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
        pb, vb = api.unproject_depth(db, cb, pipeline.MIN_DSP, pipeline.MAX_DSP, ctx)
        mt = pipeline.lattice_matches(pa, va, sc["cam_b"], vb, sc["planted"], 640, 480, seed=u)
        work.append((da, ca, db, cb, mt, sc["planted"], u))
    return work


def c4_pairs(args, api, ctx, stream, world, rank, dist, steps, warmup, clocks_index=None):
    """configs[3] (C4): 28 = C(8,2) view pairs of an 8-view turntable, each pair the C1 workload (two
    640x480 depth maps -> matched pixels unprojected -> RemoveOutliers -> SRT -> Depth2Model mesh ->
    apply -> deformation -> Model2Depth), sharded pair u -> rank u % N with no data-path collective
    (mvs_partition_units).  A step = every rank's share once, as ONE mvs_stitch_pairs call per rank
    (library worker threads, one stream each); host buffers in and out (the depth maps live in host
    memory, as the reference reads them from disk), every output of every pair written back."""
    import ctypes as C
    import torch
    from multiviewstitch_b200 import pipeline, _capi
    part = (C.c_int64 * 3)()
    _capi.check(_capi.lib().mvs_partition_units(C.c_int64(args.pairs), C.c_int(world), C.c_int(rank), part))
    mine = [part[0] + k * part[1] for k in range(part[2])]
    with torch.cuda.stream(stream):
        work = build_pair_work(api, pipeline, ctx, mine)
        batch = api.PairBatch(ctx, want_outputs=True)
        for da, ca, db, cb, mt, _, u in work:
            batch.add(da, ca, db, cb, mt, seed=u)
        cores = os.cpu_count() or 1
        want = args.pair_workers if args.pair_workers > 0 else max(1, min(16, cores // max(world, 1)))
        nw = max(1, min(want, len(work))) if work else 1
        prm = api.pair_params(n_nodes=512, gn_iters=4, workers=nw)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        outs = []
        for _ in range(warmup):
            if len(batch):
                outs = batch.run(prm)
        ctx.launch_count_reset()
        clocks = None
        if clocks_index is not None and rank == 0:
            clocks = Clocks(clocks_index)
            clocks.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        barrier()
        t0 = time.perf_counter()
        for s in range(steps):
            ev[s][0].record(stream)
            if len(batch):
                outs = batch.run(prm)  # returns after every worker stream is synchronised
            ev[s][1].record(stream)
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        launches = ctx.launch_count()
        if clocks:
            clocks.window(t0, t0 + wall)
        clk = clocks.stop() if clocks else None
    worst = 0.0
    for o, wk in zip(outs, work):  # every pair must recover its planted similarity
        ps, pR, pt = wk[5]
        worst = max(worst, abs(o["scale"] - ps) / ps, float(np.abs(o["R"] - pR).max()))
    bytes_in = sum(2 * w[0].size * 4 + w[4].size * 4 for w in work)
    bytes_out = sum(o["template"].nbytes + o["deformed"].nbytes + o["facets"].nbytes + o["render"].nbytes
                    for o in outs)
    t = torch.tensor([dev_ms, worst, float(launches)], dtype=torch.float64, device="cuda")
    tsum = torch.tensor([float(bytes_in), float(bytes_out)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    dev_ms, worst = t[0].item(), t[1].item()
    return {"metric": "view_pairs_per_s", "value": args.pairs * steps / (dev_ms * 1e-3), "unit": "pairs/s",
            "n_gpus": world, "pairs": args.pairs, "steps": steps, "warmup": warmup,
            "ms_per_step": dev_ms / steps, "scaling": "strong",
            "pairs_on_busiest_rank": -(-args.pairs // world), "host_workers_per_rank": nw,
            "host_cores": os.cpu_count(),
            "max_srt_error_vs_planted": worst, "gpu_launches_busiest_rank": int(t[2].item()),
            "h2d_bytes_per_step": int(tsum[0].item()), "d2h_bytes_per_step": int(tsum[1].item()),
            "wall_s_rank0": wall, "clocks": clk,
            "workload": f"C4: {args.pairs} view pairs (two 640x480 depth maps each) sharded pair -> rank "
                        "u % N, one mvs_stitch_pairs call per rank: RemoveOutliers + SRT + Depth2Model + "
                        "512-node deformation (4 GN x 20 PCG) + Model2Depth per pair"}


def run_pairs(args):
    """--mode pairs: configs[3] as the headline line of its own run."""
    import torch
    import torch.distributed as dist
    from multiviewstitch_b200 import api
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    ctx = api.Context(local, stream.cuda_stream)
    r = c4_pairs(args, api, ctx, stream, world, rank, dist, args.steps, max(3, args.warmup), local)
    if rank == 0:
        line = {"metric": r["metric"], "value": r["value"], "unit": r["unit"], "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32/f64",
                "data": "synthetic",
                "config": {"workload": r["workload"], "pairs_on_busiest_rank": r["pairs_on_busiest_rank"],
                           "l2": "host buffers per pair", "host_workers_per_rank": r["host_workers_per_rank"]},
                "max_srt_error_vs_planted": r["max_srt_error_vs_planted"],
                "e2e": {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": r["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": r["d2h_bytes_per_step"]},
                "gpu_launches": r["gpu_launches_busiest_rank"], "clocks": r["clocks"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------- C5: vertex-sharded graph
def c5_shard(args, api, local, stream, world, rank, dist, peak, side, nodes, steps, with_cpu=False):
    """configs[4] (C5): ONE 5M-vertex / 20K-node graph; every rank owns a contiguous range of the
    sorted vertex order (snapped to assembly groups), runs the per-vertex passes on it and the partial
    fixed-point J^T J / J^T r blocks are summed with one ncclAllReduce(int64) per GN iteration; the PCG
    runs replicated.  Integer sums: the result is BITWISE the unsharded solve — checked on rank 0 in
    the same run (shard_bitwise_equal)."""
    import torch
    from multiviewstitch_b200 import synth
    sa = argparse.Namespace(**vars(args))
    sa.side, sa.nodes = side, nodes
    verts, faces = synth.cube_sphere(side)
    tv, _ = synth.cube_sphere(int(round(side * 1.027)), rot=synth.axis_angle([1, 2, 3], 0.7))
    pts = verts.astype(np.float32)
    tgt = synth.bend_twist(tv, 0.05).astype(np.float32)
    del verts, tv
    n, nt = len(pts), len(tgt)
    ctx = api.Context(local, stream.cuda_stream)
    if world > 1:
        uid = [api.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
    out = {}
    with torch.cuda.stream(stream):
        t0 = time.perf_counter()
        d = api.Deformation(pts, None, faces, ctx)
        ids = d.SampleFPS(nodes, 0)
        d.BuildGraph(4, 8)
        stream.synchronize()
        t_setup = time.perf_counter() - t0
        st = d.Stats()
        m, nnz = st["nodes"], st["pattern_blocks"]
        if world > 1:
            d.SetShardInterleaved(rank, world, 8)  # 8 chunks per rank spread over the mesh
        tgt4 = torch.zeros((nt, 4), dtype=torch.float32, device="cuda")
        tgt4[:, :3] = torch.from_numpy(tgt).cuda()
        stream.synchronize()
        d.SetTargetDev(tgt4.data_ptr(), None, nt)
        prm = api.gn_params(gn_iters=args.gn_iters, pcg_iters=args.pcg_iters, corr_mode=1)
        flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        for _ in range(3):
            d.Solve(prm)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        barrier()
        for s in range(steps):
            flush.fill_(s & 0xFF)
            ev[s][0].record(stream)
            d.Solve(prm)
            ev[s][1].record(stream)
        barrier()
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        # one profiled solve for the per-family times of THIS rank's share
        ctx.prof_enable(True)
        ctx.prof_reset()
        d.Solve(prm)
        stream.synchronize()
        prof = ctx.prof()
        ctx.prof_enable(False)
        out_dev = d.Vertices()
        x_dev = d.NodeState()
        equal = None
        if world > 1 and rank == 0:
            ctx1 = api.Context(local, stream.cuda_stream)  # no communicator: the plain one-rank solve
            d1 = api.Deformation(pts, None, faces, ctx1)
            d1.SetNodes(ids)
            d1.BuildGraph(4, 8)
            d1.SetTargetDev(tgt4.data_ptr(), None, nt)
            d1.Solve(prm)
            equal = bool(np.array_equal(d1.Vertices(), out_dev) and np.array_equal(d1.NodeState(), x_dev))
            d1.close()
        if world > 1:
            dist.barrier()  # the ranks map each other's exchange blocks: release them together
        d.close()
        del flush, tgt4
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
    dev_ms = t[0].item()
    nup = (nnz + m) // 2
    out = {"metric": METRIC, "value": args.gn_iters * steps / (dev_ms * 1e-3), "unit": UNIT,
           "n_gpus": world, "steps": steps, "ms_per_step": dev_ms / steps, "scaling": "strong",
           "verts": n, "target_points": nt, "nodes": m, "pattern_blocks": nnz,
           "setup_s_untimed": t_setup,
           "bytes_allreduced_per_gn_iteration": int((nup * 16 + m * 12) * 8) if world > 1 else 0,
           "shard_bitwise_equal": equal,
           "families_ms_per_step_rank0": {k: round(v[0], 4) for k, v in
                                          sorted(prof.items(), key=lambda kv: -kv[1][0])},
           "workload": f"C5: {n}-vertex mesh, {m}-node graph, vertices sharded over {world} rank(s) (8 interleaved "
                       f"chunks per rank), partial fixed-point J^T J blocks + J^T r summed and converted by one "
                       f"kernel over NVLink peer memory (NCCL all-reduce as fallback) per GN iteration, "
                       f"replicated PCG; {args.gn_iters} GN x {args.pcg_iters} PCG + vertex update"}
    if with_cpu and rank == 0:
        import oracle
        t0 = time.perf_counter()
        widx, w = oracle.node_weights(pts, pts[ids], 4)
        graph = oracle.node_graph(pts[ids], 8)
        t_setup_cpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        _, oo, _ = oracle.gn_solve(pts, None, ids, widx, w, graph,
                                   oracle.gn_params(gn_iters=args.gn_iters, pcg_iters=args.pcg_iters,
                                                    corr_mode=1), tpts=tgt)
        t_cpu = time.perf_counter() - t0
        diag = float(np.linalg.norm(pts.max(0) - pts.min(0)))
        out["cpu_baseline"] = {"value": args.gn_iters / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": f"1 full solve ({t_cpu:.1f} s; node ids taken from the GPU "
                                         f"sampling, weights/graph setup {t_setup_cpu:.1f} s not counted)",
                               "max_abs_diff_vs_gpu_over_bbox_diag": float(np.abs(oo - out_dev).max() / diag)}
    ctx.close()
    return out


# ------------------------------------------------------------------------- remaining kernel rooflines
def secondary_kernels(args, api, ctx, stream, peak, d, pts, faces):
    """SURVEY §8(d) lines the headline does not cover: Model2Depth of the benchmarked mesh into a
    640x480 view (16 B/vertex + 12 B/triangle + 12 B/pixel), RANSAC hypothesis scoring on a full
    640x480 match set (32 B/match/hypothesis streamed; fp64-compute leaning: ~100 flop per match and
    hypothesis) and the PCA / bounding-box sweeps of PointSetUtils (24 B/point/pass, doubles)."""
    import torch
    from multiviewstitch_b200 import synth
    out = {}
    n, nf = len(pts), len(faces)
    with torch.cuda.stream(stream):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        # ---- Model2Depth
        K = synth.c1_intrinsics()
        cam = api.make_camera(K, np.eye(3), np.array([0.0, 0.0, 4.0]), 640, 480)
        p4 = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
        p4[:, :3] = torch.from_numpy(pts).cuda()
        f_dev = torch.from_numpy(np.ascontiguousarray(faces, np.int32)).cuda()
        img = torch.empty((480, 640), dtype=torch.float32, device="cuda")
        stream.synchronize()
        best = None
        for rep in range(4):
            ev[0].record(stream)
            api.model2depth_dev(p4.data_ptr(), n, f_dev.data_ptr(), nf, [cam], img.data_ptr(), ctx=ctx)
            ev[1].record(stream)
            stream.synchronize()
            ms = ev[0].elapsed_time(ev[1])
            best = ms if rep == 1 or (rep > 1 and ms < best) else best  # rep 0 = warm-up
        b = 16 * n + 12 * nf + 12 * 640 * 480
        out["model2depth"] = {"verts": n, "facets": nf, "image": "640x480", "ms": best,
                              "covered_pixels": int((img > 0).sum().item()),
                              "GBps": b / (best * 1e-3) / 1e9, "frac_hbm": b / (best * 1e-3) / 1e9 / peak,
                              "bytes_formula": "16 n + 12 F + 12 w h (SURVEY 8(d))",
                              "note": "includes the camera upload and the image clear of one call"}
        del p4, f_dev, img
        # ---- RANSAC scoring
        nm = 640 * 480
        (Kc, R1, t1, w, h), (_, R2, t2, _, _) = synth.camera_pair()
        c1, c2 = api.make_camera(Kc, R1, t1, w, h), api.make_camera(Kc, R2, t2, w, h)
        pp, qq = synth.planted_pairs(nm, seed=11, noise=1e-3)
        srt = api.SRTSolver(200, ctx)
        srt.SetPrintFlag(False)
        srt.SetInput(pp, qq, c1, c2)
        srt.EstimateTransformRansac(None, seed=1)
        ctx.prof_enable(True)
        ctx.prof_reset()
        reps = 3
        for rep in range(reps):
            srt.EstimateTransformRansac(None, seed=2 + rep)
        prof = ctx.prof()
        ctx.prof_enable(False)
        srt.close()
        ms = prof.get("srt_residual", (0.0, 0))[0] / reps
        if ms > 0:
            hm = 200.0 * nm
            out["srt_ransac_scoring"] = {"matches": nm, "hypotheses": 200, "ms": ms,
                                         "hypothesis_matches_per_s": hm / (ms * 1e-3),
                                         "GBps_streamed_model": 48.0 * hm / (ms * 1e-3) / 1e9,
                                         "frac_hbm_streamed_model": 48.0 * hm / (ms * 1e-3) / 1e9 / peak,
                                         "fp64_GFLOPs": 100.0 * hm / (ms * 1e-3) / 1e9,
                                         "note": "matches are six float64 planes (48 B/match) that stay in "
                                                 "L2 across the 200 hypotheses; the kernel is fp64-compute "
                                                 "bound (4 projections = ~100 fp64 flop + 8 divisions per "
                                                 "match and hypothesis), the GB/s is the streamed-model figure"}
        # ---- PointSetUtils sweeps
        np_pts = 4_000_000
        big = np.random.default_rng(3).normal(size=(np_pts, 3))
        psu = api.PointSetUtils(ctx)
        psu.SetInput(big)
        psu.CalcPivots(3)
        ctx.prof_enable(True)
        ctx.prof_reset()
        psu.SetInput(big)
        psu.CalcPivots(3)
        prof = ctx.prof()
        ctx.prof_enable(False)
        sweeps = {}
        for k in ("ps_sum", "ps_bbox", "ps_cov"):
            if k in prof and prof[k][0] > 0:
                msk = prof[k][0] / max(prof[k][1], 1)
                sweeps[k] = {"ms": msk, "GBps": 24.0 * np_pts / (msk * 1e-3) / 1e9,
                             "frac_hbm": 24.0 * np_pts / (msk * 1e-3) / 1e9 / peak}
        out["pointset_sweeps"] = {"points": np_pts, "bytes_per_point_per_pass": 24, "kernels": sweeps}
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from multiviewstitch_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    ctx = api.Context(local, stream.cuda_stream)
    shard = args.mode == "shard" and world > 1

    peaks = {}
    pk_file = ROOT / "MEASURED_PEAKS.json"
    if pk_file.exists():
        peaks = json.loads(pk_file.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))

    # every rank solves ITS OWN copy of the same object (weak scaling: the per-GPU work is fixed; a
    # different similarity per rank changes how the scan falls into search cells and with it the
    # correspondence time by several per cent — that is object variance, not scaling)
    pts, nrm, faces, tgt = workload(args, 0)
    n, nt = len(pts), len(tgt)
    if shard:
        uid = [api.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        d = api.Deformation(pts, nrm, faces, ctx)
        d.SampleFPS(args.nodes, 0)
        d.BuildGraph(4, 8)
        m = len(d.Nodes())
        st = d.Stats()
        nnz = st["pattern_blocks"]
        global NRUNS
        NRUNS = int(st["tuple_runs"])
        if shard:
            d.SetShardInterleaved(rank, world, 8)
        # device-resident target (float4 records) for the HBM-resident leg
        tgt4 = torch.zeros((nt, 4), dtype=torch.float32, device="cuda")
        tgt4[:, :3] = torch.from_numpy(tgt).cuda()
        stream.synchronize()
        d.SetTargetDev(tgt4.data_ptr(), None, nt)
        prm = api.gn_params(gn_iters=args.gn_iters, pcg_iters=args.pcg_iters, corr_mode=1)
        flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

        clocks = Clocks(local)
        if rank == 0:
            clocks.start()  # polling runs through the warm-up; the samples inside the timed region count
        for _ in range(max(3, args.warmup)):
            d.Solve(prm)
        # ---- timed: device-resident, one library call per step, profiling off
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(args.steps)]
        ctx.launch_count_reset()
        barrier()
        t_wall0 = time.perf_counter()
        for s in range(args.steps):
            flush.fill_(s & 0xFF)  # L2 flush between timed steps (outside the event pair)
            ev[s][0].record(stream)
            d.Solve(prm)
            ev[s][1].record(stream)
        barrier()
        t_wall = time.perf_counter() - t_wall0
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        launches = ctx.launch_count()
        clocks.window(t_wall0, t_wall0 + t_wall)
        clk = clocks.stop() if rank == 0 else None
        out_dev = d.Vertices()
        x_dev = d.NodeState()

        # ---- the same K steps again with the per-family event profiling on (families, roofline)
        ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(args.steps)]
        ctx.prof_enable(True)
        ctx.prof_reset()
        for s in range(args.steps):
            flush.fill_(s & 0xFF)
            ev3[s][0].record(stream)
            d.Solve(prm)
            ev3[s][1].record(stream)
        stream.synchronize()
        prof = ctx.prof()
        ctx.prof_enable(False)
        prof_ms = sum(a.elapsed_time(b) for a, b in ev3)

        # ---- timed: end to end through the host-buffer C ABI
        e2e_ms = None
        if not args.no_e2e and not shard:
            tgt_pin = torch.from_numpy(tgt).pin_memory()
            out_pin = torch.empty((n, 3), dtype=torch.float32).pin_memory()
            tgt_h, out_h = tgt_pin.numpy(), out_pin.numpy()
            import ctypes as C

            def solve_e2e():
                d.SetTarget(tgt_h)  # H2D + NN structure build
                d.Solve(prm)
                api.check(api.lib().mvs_deform_get_vertices_f32(
                    d._h, out_h.ctypes.data_as(C.POINTER(C.c_float))))  # D2H

            for _ in range(2):
                solve_e2e()
            ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                   for _ in range(args.steps)]
            barrier()
            for s in range(args.steps):
                flush.fill_(s & 0xFF)
                ev2[s][0].record(stream)
                solve_e2e()
                ev2[s][1].record(stream)
            barrier()
            e2e_ms = sum(a.elapsed_time(b) for a, b in ev2)
            # fixed-point block sums: the two legs agree bit for bit
            dg = float(np.linalg.norm(pts.max(0) - pts.min(0)))
            e2e_diff = float(np.abs(out_h - out_dev).max() / dg)
            d.SetTargetDev(tgt4.data_ptr(), None, nt)

    # ---- vertex-sharded mode: the all-reduced solve must carry the bits of the unsharded one
    shard_equal = None
    if shard:
        if rank == 0:
            with torch.cuda.stream(stream):
                ctx1 = api.Context(local, stream.cuda_stream)  # no communicator: plain one-rank solve
                d1 = api.Deformation(pts, nrm, faces, ctx1)
                d1.SampleFPS(args.nodes, 0)
                d1.BuildGraph(4, 8)
                d1.SetTargetDev(tgt4.data_ptr(), None, nt)
                d1.Solve(prm)
                shard_equal = bool(np.array_equal(d1.Vertices(), out_dev)
                                   and np.array_equal(d1.NodeState(), x_dev))
                d1.close()
        dist.barrier()

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms or 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_max = t[0].item(), t[1].item()
        if e2e_ms is not None:
            e2e_ms = e2e_max
    solves_total = args.steps * (1 if shard else world)
    value = args.gn_iters * solves_total / (dev_ms * 1e-3)

    line = None
    if rank == 0:
        # roofline of the dominant kernel family (profiled pass over the same K steps)
        fam = {k: v for k, v in prof.items() if algo_bytes(k, n, nt, m, nnz, args.pcg_iters)}
        top = max(fam, key=lambda k: fam[k][0])
        avg_ms = fam[top][0] / max(fam[top][1], 1)
        n_local = n
        if shard:
            per = (n + world - 1) // world
            n_local = min(n, (rank + 1) * per) - rank * per
        ab = algo_bytes(top, n_local, nt, m, nnz, args.pcg_iters)
        achieved = ab / (avg_ms * 1e-3) / 1e9
        traffic = None
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists():
            traffic = json.loads(tf.read_text()).get(top)
        roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "algorithmic_bytes_per_launch": ab, "avg_ms": avg_ms,
                    "launches_timed": int(fam[top][1]),
                    "bytes_formula": ALGO_BYTES_DOC.get(top),
                    "measured": "CUDA events inside the library around every launch of the family, "
                                "profiled pass over the same K steps (the timed pass runs with them off)"}
        families = {}
        for k, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            b = algo_bytes(k, n_local, nt, m, nnz, args.pcg_iters)
            families[k] = {"ms_per_step": ms / args.steps, "launches": cnt, "share": ms / prof_ms}
            if b and cnt:
                families[k]["GBps"] = b / (ms / cnt * 1e-3) / 1e9
                families[k]["frac_hbm"] = families[k]["GBps"] / peak

        cpu = None
        if not args.no_cpu_baseline and world == 1:
            ref_it = args.ref_gn_iters if args.ref_gn_iters > 0 else args.gn_iters
            t_setup, t_solve, out_cpu, graph_parts = cpu_solve(args, pts, tgt, ref_it)
            diag = float(np.linalg.norm(pts.max(0) - pts.min(0)))
            cpu = {"value": ref_it / t_solve, "unit": UNIT, "cores": 1, "kind": "port",
                   "host_cores": os.cpu_count(),
                   "sample": f"1 solve of {ref_it} GN x {args.pcg_iters} PCG iterations on the "
                             f"full workload ({t_solve:.1f} s; graph setup {t_setup:.1f} s not counted)"}
            if ref_it == args.gn_iters:
                cpu["max_abs_diff_vs_gpu_over_bbox_diag"] = float(np.abs(out_cpu - out_dev).max() / diag)
            try:
                cpu["openmp_courtesy"] = cpu_solve_openmp(args, pts, tgt, ref_it, graph_parts)
            except Exception as exc:  # never lose the headline line over the courtesy number
                cpu["openmp_courtesy"] = {"error": repr(exc)}

        secondary = {}
        if not args.no_secondary and world == 1:
            try:
                secondary.update(secondary_c2(args, api, ctx, stream, peak))
            except Exception as exc:  # the headline line must still be printed
                secondary["c2_error"] = repr(exc)
            try:
                secondary["arap"] = secondary_arap(args, api, ctx, stream, peak, d, pts, faces)
            except Exception as exc:
                secondary["arap"] = {"error": repr(exc)}
            try:
                secondary["kernels"] = secondary_kernels(args, api, ctx, stream, peak, d, pts, faces)
            except Exception as exc:
                secondary["kernels"] = {"error": repr(exc)}

        cfg = base_config(args, n, nt)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "strong" if shard else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg,
                "graph": {"pattern_blocks": nnz, "tuple_runs": st["tuple_runs"]},
                "roofline": roofline, "cpu_baseline": cpu,
                "e2e": None if e2e_ms is None else {
                    "value": args.gn_iters * args.steps * world / (e2e_ms * 1e-3), "unit": UNIT,
                    "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(nt * 12), "d2h_bytes_per_step": int(n * 12),
                    "max_abs_diff_vs_resident_over_bbox_diag": e2e_diff},
                "gpu_launches": int(launches), "clocks": clk,
                "wall_s_timed_region": t_wall, "profiled_pass_ms_per_step": prof_ms / args.steps,
                "secondary": secondary, "families": families,
                "shard_bitwise_equal": shard_equal}

    # ---- release the headline workload before the two partitions
    with torch.cuda.stream(stream):
        if world > 1:
            dist.barrier()  # sharded handles map each other's exchange blocks: release them together
        d.close()
        del tgt4, flush
    torch.cuda.empty_cache()

    # ---- north_star's partitions, at every N (secondary.c4_pairs, secondary.c5_shard)
    if not args.no_secondary and args.mode == "objects":
        c4 = c5 = None
        if not args.no_c4:
            try:
                c4 = c4_pairs(args, api, ctx, stream, world, rank, dist, 3, 3)
            except Exception as exc:
                c4 = {"error": repr(exc)}
        if not args.no_c5:
            try:
                c5 = c5_shard(args, api, local, stream, world, rank, dist, peak, args.c5_side,
                              args.c5_nodes, args.c5_steps, with_cpu=args.c5_cpu)
            except Exception as exc:
                c5 = {"error": repr(exc)}
        if rank == 0:
            line["secondary"]["c4_pairs"] = c4
            line["secondary"]["c5_shard"] = c5
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "pairs":
        run_pairs(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
