"""Position sharding across the GPUs of one box.

The pileup text is cut into contiguous byte ranges, one per rank.  A rank owns the lines whose FIRST
byte lies in its range (the tokenizer kernel applies the rule itself, so ranges may cut lines
anywhere); `local` and `quality` need no communication at all.  Methods with a Lynch fit exchange
their unique-profile histograms ONCE: an NCCL all-gather of (profile, count) pairs, a few thousand per
rank.  Every rank merges them on its device (sidgpu_set_global_histogram: counts summed, lexicographic
order = countUniqueProfiles of the whole genome) and runs the whole Nelder-Mead fit as one kernel on
the merged histogram: identical input, identical code, so (pi, eps) are bit-identical on every rank and
for every number of ranks, with no traffic per optimiser step.  likelihood_ratio's Benjamini-Hochberg
ranks (m = unique profiles of the whole genome) come from the same merged list.

The literal north_star form -- each rank reduces its own histogram and the ranks all-reduce ONE double
per objective evaluation (lynch.cpp:37-61) -- is kept as `per_evaluation_allreduce=True` (and as the
gloo-tested distributed_fit); bench.py times both."""
from . import nelder_mead


def shard_ranges(text_len, world):
    """Contiguous byte ranges [begin, end) of a text of text_len bytes, one per rank."""
    base, rem = divmod(text_len, world)
    out, pos = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((pos, pos + n))
        pos += n
    return out


def owned_line_starts(text, begin, end):
    """Host restatement of the ownership rule for tests: offsets p in [begin, end) with
    text[p] != '\\n' and (p == 0 or text[p-1] == '\\n')."""
    out = []
    for p in range(begin, end):
        if text[p] != 10 and (p == 0 or text[p - 1] == 10):
            out.append(p)
    return out


def distributed_fit(local_sums, local_objective, all_reduce_ints, all_reduce_float):
    """Lynch fit over sharded histograms.
      local_sums()                 -> [A, C, G, T, total] integer sums of this rank's histogram
      local_objective(nd, pi, eps) -> this rank's partial of the objective (a float, or an object the
                                      all-reduce understands)
      all_reduce_ints(list)        -> element-wise sum over ranks
      all_reduce_float(x)          -> sum over ranks, as a Python float
    Returns dict(pi, eps, nd, fval, iterations, evaluations, converged); identical on every rank."""
    sums = all_reduce_ints(list(local_sums()))
    if sums[4]:
        nd = [sums[i] / sums[4] for i in range(4)]          # pileup.cpp:209-213
    else:
        nd = [0.25] * 4                                     # pileup.cpp:215

    def f(pi, eps):
        if pi < 0 or pi > 1 or eps < 0 or eps > 1:          # lynch.cpp:41-43, decided before any exchange
            return 1.7976931348623157e308
        return all_reduce_float(local_objective(nd, pi, eps))

    r = nelder_mead.nelder_mead_2d(f)
    return {"pi": r["x"][0], "eps": r["x"][1], "nd": nd, "fval": r["fval"], "iterations": r["iterations"],
            "evaluations": r["evaluations"], "converged": r["converged"]}


def profile_sort_key(p):
    """Lexicographic order of std::array<uint16_t,4> (pileup.cpp:179-182) for packed profiles (numpy uint64)."""
    import numpy as np
    p = np.asarray(p, dtype=np.uint64)
    m = np.uint64(0xFFFF)
    return ((p & m) << np.uint64(48)) | (((p >> np.uint64(16)) & m) << np.uint64(32)) | (((p >> np.uint64(32)) & m) << np.uint64(16)) | (p >> np.uint64(48))


def merge_histograms(tables):
    """Merges per-rank (profiles, counts) histograms: one entry per profile, counts summed, in the
    reference's lexicographic order (countUniqueProfiles over the whole genome)."""
    import numpy as np
    prof = np.concatenate([np.asarray(t[0], dtype=np.uint64) for t in tables]) if tables else np.zeros(0, np.uint64)
    cnt = np.concatenate([np.asarray(t[1], dtype=np.uint64) for t in tables]) if tables else np.zeros(0, np.uint64)
    if prof.size == 0:
        return prof, cnt
    u, inv = np.unique(prof, return_inverse=True)
    c = np.zeros(len(u), dtype=np.uint64)
    np.add.at(c, inv, cnt)
    order = np.argsort(profile_sort_key(u), kind="stable")
    return u[order], c[order]


def torch_collectives(dist, device):
    """(all_reduce_ints, all_reduce_float) over torch.distributed for distributed_fit."""
    import torch

    def ints(v):
        t = torch.tensor(v, dtype=torch.int64, device=device)
        dist.all_reduce(t)
        return [int(x) for x in t.tolist()]

    def flt(x):
        t = x if isinstance(x, torch.Tensor) else torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t)
        return float(t.reshape(-1)[0].item())

    return ints, flt


def all_gather_histograms(dist, prof, cnt):
    """All-gather of every rank's (profiles, counts) histogram (int64 tensors of any, rank-dependent, length on the
    collective's device).  Returns (profiles, counts) of world * max_length entries; the padding has count 0."""
    import torch
    world = dist.get_world_size()
    n = torch.tensor([prof.numel()], dtype=torch.int64, device=prof.device)
    dist.all_reduce(n, op=dist.ReduceOp.MAX)
    m = max(1, int(n.item()))
    p = torch.zeros(m, dtype=torch.int64, device=prof.device)
    c = torch.zeros(m, dtype=torch.int64, device=prof.device)
    p[:prof.numel()] = prof
    c[:cnt.numel()] = cnt
    gp = torch.empty(world * m, dtype=torch.int64, device=prof.device)
    gc = torch.empty(world * m, dtype=torch.int64, device=prof.device)
    dist.all_gather_into_tensor(gp, p)
    dist.all_gather_into_tensor(gc, c)
    return gp, gc


_exchange_buffers = {}


def exchange_histograms(ctx, dist, device, shared_stream=False):
    """The collective step of a sharded session with a genome-wide fit: ONE all-gather of the ranks' histograms
    (NCCL; profiles and counts in one buffer), merged on this rank's device.  Returns the number of gathered entries.
    shared_stream: the ctx was created on torch's current stream, so its kernels, the copies and the collective are
    already in order and nothing has to be synchronised in between."""
    import torch
    world = dist.get_world_size()
    n_u, d_prof, d_cnt = ctx.histogram_device(4)
    n = torch.tensor([n_u], dtype=torch.int64, device=device)
    dist.all_reduce(n, op=dist.ReduceOp.MAX)
    m = max(1, int(n.item()))
    cap = 1 << (m - 1).bit_length()                         # buffers are kept and only grow
    key = (str(device), world)
    buf = _exchange_buffers.get(key)
    if buf is None or buf[0].numel() < 2 * cap:
        buf = (torch.zeros(2 * cap, dtype=torch.int64, device=device), torch.empty(world * 2 * cap, dtype=torch.int64, device=device))
        _exchange_buffers[key] = buf
    mine, gathered = buf
    cap = mine.numel() // 2
    # layout per rank: cap profiles, then cap counts; a count of 0 marks padding
    mine[cap:].zero_()
    if not shared_stream:
        torch.cuda.current_stream().synchronize()           # the clearing ran on torch's stream, the copies run on the ctx's
    if n_u:
        ctx.copy_d2d(mine.data_ptr(), d_prof, 8 * n_u, sync=not shared_stream)
        ctx.copy_d2d(mine.data_ptr() + 8 * cap, d_cnt, 8 * n_u, sync=not shared_stream)
    dist.all_gather_into_tensor(gathered, mine)
    if not shared_stream:
        torch.cuda.current_stream().synchronize()           # the collective ran on torch's stream, the merge runs on the ctx's
    g = gathered.view(world, 2, cap)
    prof = g[:, 0, :].contiguous()
    cnt = g[:, 1, :].contiguous()
    ctx.set_global_histogram(prof.data_ptr(), cnt.data_ptr(), world * cap)
    _exchange_buffers[key + ("keep",)] = (prof, cnt)        # alive until the merge kernels have run
    return world * cap


def call_sharded(ctx, d_text, text_len, rank, world, params, dist=None, device=None, per_evaluation_allreduce=False):
    """Runs one calling session on this rank's shard of a device-resident text that every rank
    holds (or at least its own range plus the straddling line).  Returns (n_sites, fit or None);
    rows are then available through ctx.emit_csv(0, n_sites, ...)."""
    import torch
    begin, end = shard_ranges(text_len, world)[rank]
    ctx.begin(params)
    n = ctx.feed(d_text, text_len, begin, end)
    needs_fit = params.method in (1, 2) or params.estimate_prior
    fit = None
    if needs_fit and world > 1 and not params.fit_given and not per_evaluation_allreduce:
        exchange_histograms(ctx, dist, device)
        ctx.finish()
        return n, ctx.session_fit()
    if needs_fit and world > 1 and not params.fit_given:
        ints, flt = torch_collectives(dist, device)
        obj = torch.zeros(1, dtype=torch.float64, device=device)
        _, sums = ctx.histogram_sums(4)

        def local_objective(nd, pi, eps):
            ctx.lynch_objective_partial(nd, pi, eps, obj.data_ptr())
            ctx.synchronize()       # the kernel ran on the ctx's stream, the all-reduce runs on torch's: order them
            return obj

        fit = distributed_fit(lambda: sums, local_objective, ints, flt)
        ctx.set_fit(fit["pi"], fit["eps"], fit["nd"])
        if params.method == 2:
            # Benjamini-Hochberg ranks over the unique profiles of the whole genome: gather the shards'
            # histograms (a few thousand entries each), merge on the host, finish on the device
            n_u, d_prof, d_cnt = ctx.histogram_device(4)
            prof = torch.empty(n_u, dtype=torch.int64, device=device)
            cnt = torch.empty(n_u, dtype=torch.int64, device=device)
            ctx.copy_d2d(prof.data_ptr(), d_prof, 8 * n_u)
            ctx.copy_d2d(cnt.data_ptr(), d_cnt, 8 * n_u)
            gp, gc = all_gather_histograms(dist, prof, cnt)
            keep = gc.cpu().numpy() > 0
            merged, _ = merge_histograms([(gp.cpu().numpy().view("uint64")[keep], gc.cpu().numpy().view("uint64")[keep])])
            ctx.finish_global(merged)
            return n, fit
    ctx.finish()
    if needs_fit and fit is None:
        fit = ctx.session_fit()
    return n, fit
