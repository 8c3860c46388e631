"""GPU: the sharded path on one device -- two contexts play two ranks (ranges of the same text), the
"all-reduce" is a sum over the two contexts.  Checks sidgpu_set_fit / histogram sums / the partial
objective and that concatenated shard output equals the single-context output."""
import numpy as np
import pytest

import oracle_py as op
from sid_b200 import shard
from test_oracle import read

pytestmark = pytest.mark.gpu


def _emit(ctx, n):
    cap = max(4096, n * 80)
    buf = ctx.device_buffer(cap)
    try:
        nbytes, rows = ctx.emit_csv(0, n, buf, cap)
        return buf.download(np.uint8, nbytes).tobytes(), rows
    finally:
        buf.free()


@pytest.mark.parametrize("method", ["local", "bayes", "likelihood_ratio"])
def test_two_shards_equal_one(native, gpu_ctx, method):
    import sid_b200
    text = read("depth30_two_chroms.plp")
    params = sid_b200.Context.make_params(method)
    whole_rows, _, _ = gpu_ctx.call_host(text, params)
    whole_fit = gpu_ctx.session_fit() if method != "local" else None
    ctxs = [sid_b200.Context(), sid_b200.Context()]
    bufs = [c.upload_text(text) for c in ctxs]
    try:
        ranges = shard.shard_ranges(len(text), 2)
        ns = []
        for c, d, (b, e) in zip(ctxs, bufs, ranges):
            c.begin(sid_b200.Context.make_params(method))
            ns.append(c.feed(d, len(text), b, e))
        if method != "local":
            sums = [c.histogram_sums(4)[1] for c in ctxs]
            fit = shard.distributed_fit(lambda: [sum(s[i] for s in sums) for i in range(5)],
                                        lambda nd, pi, eps: sum(c.lynch_objective(nd, pi, eps) for c in ctxs),
                                        lambda v: v, lambda x: x)
            assert fit["converged"]
            assert abs(fit["pi"] - whole_fit["pi"]) <= 1e-6 * whole_fit["pi"]
            assert abs(fit["eps"] - whole_fit["eps"]) <= 1e-6 * whole_fit["eps"]
            assert np.allclose(fit["nd"], whole_fit["nd"], rtol=0, atol=1e-15)
            for c in ctxs:
                c.set_fit(fit["pi"], fit["eps"], fit["nd"])
        merged = None
        if method == "likelihood_ratio":
            merged, counts = shard.merge_histograms([c.histogram(4)[:2] for c in ctxs])
            o = op.oracle_call(text, "likelihood_ratio")
            prof = o["profiles"]
            cov = op.unpack_profiles(prof).astype(np.int64).sum(axis=1)
            wu, wc = op.oracle_unique(prof[cov >= 4])
            assert np.array_equal(merged, wu) and np.array_equal(counts, wc)      # == countUniqueProfiles of the whole text
        rows = b""
        for c, n in zip(ctxs, ns):
            if merged is not None:
                c.finish_global(merged)
            else:
                c.finish()
            r, _ = _emit(c, n)
            rows += r
        n, diffs = op.compare_csv(rows, whole_rows)
        assert diffs <= 2
    finally:
        for b in bufs:
            b.free()
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("method", ["bayes", "likelihood_ratio"])
@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_merged_histograms_give_the_single_gpu_fit_bit_for_bit(native, gpu_ctx, method, n_shards):
    """The collective form (sidgpu_set_global_histogram): every shard gets the histograms of all shards (here the
    "all-gather" is a concatenation), merges them on the device and runs the whole fit there.  The merged histogram is
    countUniqueProfiles of the whole text, so (pi, eps, iterations) are EQUAL to the single-GPU fit for any number of
    shards, and the concatenated rows are the single-GPU rows."""
    import sid_b200
    text = read("depth30_two_chroms.plp")
    whole_rows, _, _ = gpu_ctx.call_host(text, sid_b200.Context.make_params(method))
    whole_fit = gpu_ctx.session_fit()
    ctxs = [sid_b200.Context() for _ in range(n_shards)]
    bufs = [c.upload_text(text) for c in ctxs]
    try:
        ns = []
        for c, d, (b, e) in zip(ctxs, bufs, shard.shard_ranges(len(text), n_shards)):
            c.begin(sid_b200.Context.make_params(method))
            ns.append(c.feed(d, len(text), b, e))
        hists = [c.histogram(4)[:2] for c in ctxs]
        m = max(1, max(len(h[0]) for h in hists))
        prof = np.zeros(n_shards * m, dtype=np.uint64)
        cnt = np.zeros(n_shards * m, dtype=np.uint64)                  # padding: count 0, as the all-gather pads
        for k, (p, c) in enumerate(hists):
            prof[k * m:k * m + len(p)] = p
            cnt[k * m:k * m + len(c)] = c
        rows = b""
        for c, n in zip(ctxs, ns):
            dp, dc = c.device_buffer(prof.nbytes).upload(prof), c.device_buffer(cnt.nbytes).upload(cnt)
            try:
                c.set_global_histogram(dp.ptr, dc.ptr, len(prof))
                c.finish()
            finally:
                dp.free()
                dc.free()
            f = c.session_fit()
            assert (f["pi"], f["eps"], f["iterations"], f["evaluations"]) == (whole_fit["pi"], whole_fit["eps"], whole_fit["iterations"], whole_fit["evaluations"])
            assert f["nd"] == whole_fit["nd"] and f["n_unique"] == whole_fit["n_unique"]
            r, _ = _emit(c, n)
            rows += r
        assert rows == whole_rows
    finally:
        for b in bufs:
            b.free()
        for c in ctxs:
            c.close()
