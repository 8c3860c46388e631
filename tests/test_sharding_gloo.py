"""CPU: the N>1 host logic with two gloo ranks -- byte-range sharding with the line-ownership rule,
and the distributed Lynch fit (one all-reduce per optimiser evaluation).  The per-rank objective
here is the host build of the device arithmetic (tests/hostcheck); on the GPU box the same driver
runs on the CUDA reduction (test_gpu_multi.py, bench.py --gpus N)."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_py as op
from sid_b200 import nelder_mead, shard
from test_oracle import read


def test_shard_ranges_cover_and_ownership():
    text = read("depth30_two_chroms.plp")
    whole = shard.owned_line_starts(text, 0, len(text))
    for world in (1, 2, 3, 4, 8):
        ranges = shard.shard_ranges(len(text), world)
        assert ranges[0][0] == 0 and ranges[-1][1] == len(text)
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        got = []
        for b, e in ranges:
            got += shard.owned_line_starts(text, b, e)
        assert got == whole                      # every line owned exactly once, in order


def test_python_nelder_mead_equals_oracle(native):
    """The Python driver walks the same trajectory as the oracle's restatement."""
    text = read("depth30.plp")
    o = op.oracle_call(text, "bayes")
    prof = o["profiles"]
    cov = op.unpack_profiles(prof).astype(np.int64).sum(axis=1)
    u, c = op.oracle_unique(prof[cov >= 4])
    nd = op.oracle_nd(u, c)
    r = nelder_mead.nelder_mead_2d(lambda pi, eps: op.oracle_objective(u, c, nd, pi, eps))
    assert r["iterations"] == o["iterations"] and r["evaluations"] == o["evaluations"]
    assert r["x"][0] == o["pi"] and r["x"][1] == o["eps"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, text, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hc = op.hostcheck()
        begin, end = shard.shard_ranges(len(text), world)[rank]
        starts = shard.owned_line_starts(text, begin, end)
        # this rank's sites: tokenised with the host build of the device tokenizer
        prof = []
        hl = op.HcLine()
        for s in starts:
            hc.hc_parse_line(text, len(text), s, 0, ctypes.byref(hl))
            assert hl.status == 0
            prof.append(hl.profile)
        prof = np.array(prof, dtype=np.uint64)
        cov = op.unpack_profiles(prof).astype(np.int64).sum(axis=1)
        u, c = np.unique(prof[cov >= 4], return_counts=True)
        u = np.ascontiguousarray(u)
        c = np.ascontiguousarray(c.astype(np.uint64))
        p4 = op.unpack_profiles(u).astype(np.int64)
        sums = [int((p4[:, i] * c.astype(np.int64)).sum()) for i in range(4)] + [int((p4.sum(axis=1) * c.astype(np.int64)).sum())]

        def local_objective(nd, pi, eps):
            a = (ctypes.c_double * 4)(*nd)
            return hc.hc_lynch_objective(len(u), u.ctypes.data, c.ctypes.data, a, pi, eps)

        ints, flt = shard.torch_collectives(dist, "cpu")
        fit = shard.distributed_fit(lambda: sums, local_objective, ints, flt)
        out.put((rank, len(starts), fit))
    finally:
        dist.destroy_process_group()


def test_distributed_fit_two_ranks(native):
    text = read("depth30.plp")
    want = op.oracle_call(text, "bayes")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, text, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] + res[1][1] == want["n_sites"]            # the two shards partition the lines
    f0, f1 = res[0][2], res[1][2]
    assert f0 == f1                                             # both ranks walked the identical trajectory
    assert f0["converged"]
    assert abs(f0["pi"] - want["pi"]) <= 1e-6 * want["pi"] and abs(f0["eps"] - want["eps"]) <= 1e-6 * want["eps"]
    assert f0["iterations"] == want["iterations"]


def test_merge_histograms_equals_count_unique(native):
    """Merging per-shard histograms gives countUniqueProfiles of the whole text (order included)."""
    text = read("depth30.plp")
    o = op.oracle_call(text, "local")
    prof = o["profiles"]
    wu, wc = op.oracle_unique(prof)
    parts = np.array_split(prof, 3)
    tables = []
    for part in parts:
        u, c = np.unique(part, return_counts=True)
        tables.append((u, c.astype(np.uint64)))
    mu, mc = shard.merge_histograms(tables)
    assert np.array_equal(mu, wu) and np.array_equal(mc, wc)
