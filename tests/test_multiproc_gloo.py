"""World-size-2 tests of the N>1 host logic on CPU (gloo): the unit partition of the pair-sharded
mode and the linearity the vertex-sharded Gauss-Newton step relies on — partial J^T J / J^T r of
disjoint vertex shards, all-reduced, plus the node-only terms, equal the full normal equations.
(On the GPU box the same exchange is one ncclAllReduce inside mvs_deform_gn_step.)"""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from tests import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partition(n_units, nranks, rank):
    from multiviewstitch_b200 import _capi
    out = (C.c_int64 * 3)()
    _capi.check(_capi.lib().mvs_partition_units(C.c_int64(n_units), C.c_int(nranks), C.c_int(rank), out))
    begin, stride, count = out[0], out[1], out[2]
    return [begin + k * stride for k in range(count)]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- pair sharding: 28 view pairs over the ranks (Processor.cpp:629-826 independent pairs)
        mine = _partition(28, world, rank)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        # ---- vertex-sharded assembly
        verts, faces = synth.uv_sphere(24, 16)
        pts = verts.astype(np.float32)
        ids = oracle.sample_fps(pts, 12, 0)
        widx, w = oracle.node_weights(pts, pts[ids], 4)
        graph = oracle.node_graph(pts[ids], 8)
        targets = synth.bend_twist(verts, 0.05)
        rng = np.random.default_rng(1)
        x = np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0]), (12, 1)) + 0.03 * rng.normal(size=(12, 12))
        n = len(pts)
        per = (n + world - 1) // world
        lo, hi = rank * per, min(n, (rank + 1) * per)
        tw = np.zeros(n, np.float32)
        tw[lo:hi] = 1.0
        fit_only = oracle.gn_params(w_rot=0.0, w_reg=0.0, lam=0.0)
        H, g = oracle.gn_assemble_dense(pts, ids, widx, w, graph, targets, tw, fit_only, x)
        Ht, gt = torch.from_numpy(H), torch.from_numpy(g)
        dist.all_reduce(Ht)
        dist.all_reduce(gt)
        node_only = oracle.gn_params(w_fit=0.0, lam=0.0)
        Hn, gn = oracle.gn_assemble_dense(pts, ids, widx, w, graph, targets, np.zeros(n, np.float32), node_only, x)
        Hf, gf = oracle.gn_assemble_dense(pts, ids, widx, w, graph, targets, None, oracle.gn_params(lam=0.0), x)
        ok = (np.allclose(Ht.numpy() + Hn, Hf, rtol=1e-12, atol=1e-9 * np.abs(Hf).max())
              and np.allclose(gt.numpy() + gn, gf, rtol=1e-12, atol=1e-9 * np.abs(gf).max()))
        # ---- max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            q.put((gathered, bool(ok), float(t.item())))
    finally:
        dist.destroy_process_group()


def test_world_size_2_partition_and_sharded_assembly():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, _free_port() if r == 0 else 0, q)) for r in range(world)]
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    gathered, ok, tmax = q.get()
    allu = sorted(u for part in gathered for u in part)
    assert allu == list(range(28)) and abs(len(gathered[0]) - len(gathered[1])) <= 1
    assert ok
    assert tmax == 2.0


def test_partition_units_edge_cases():
    for n_units, nranks in ((0, 4), (3, 8), (28, 8), (28, 1), (29, 4)):
        parts = [_partition(n_units, nranks, r) for r in range(nranks)]
        assert sorted(u for p in parts for u in p) == list(range(n_units))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
