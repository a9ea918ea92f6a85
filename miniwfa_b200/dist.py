"""Multi-GPU batches: independent pairs are sharded across ranks (one process per GPU), every rank aligns its shard on
its own device with no data-path collective, and ONE gather at the end brings the mwf_rst_t records to rank 0
(the reference has no counterpart -- its CLI loops over pairs, main.c:67).

The gather payload per pair is the fixed header {s, n_cigar, n_iter} followed by the CIGAR words, flattened into one
int64 tensor per rank and padded to the largest rank's size, so a single `gather` suffices (NCCL over NVLink on the
GPU box, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_pairs, world, rank, costs=None):
    """Pairs of this rank.  Default: round-robin (i mod world == rank).  With per-pair cost estimates, greedy
    longest-processing-time assignment, so that ragged batches balance; ties keep the input order."""
    if costs is None:
        return list(range(rank, n_pairs, world))
    order = sorted(range(n_pairs), key=lambda i: (-costs[i], i))
    load = [0] * world
    mine = []
    for i in order:
        g = min(range(world), key=lambda r: (load[r], r))
        load[g] += costs[i]
        if g == rank:
            mine.append(i)
    return sorted(mine)


def pair_cost(t, q):
    """Work estimate for balancing: wavefront cells grow with the square of the score, and the score with the
    length difference plus the divergence; without knowing the divergence, length is the usable proxy."""
    n = max(len(t), len(q))
    return n * n + 1


def pack_results(idx, results):
    """[(s, n_cigar, n_iter, [cigar words])] of pairs `idx` -> one flat int64 array: n, then per pair idx, s, n_cigar, n_iter and
    its CIGAR words (numpy concatenation: no Python-level loop over the words)."""
    parts = [np.array([len(idx)], dtype=np.int64)]
    for i, r in zip(idx, results):
        parts.append(np.array([i, r[0], r[1], r[2]], dtype=np.int64))
        if r[1] > 0:
            parts.append(np.asarray(r[3], dtype=np.int64))
    return np.concatenate(parts)


def unpack_results(flat, out):
    flat = np.asarray(flat, dtype=np.int64)
    n = int(flat[0])
    p = 1
    for _ in range(n):
        i, s, nc, ni = (int(x) for x in flat[p:p + 4])
        p += 4
        out[i] = (s, nc, ni, flat[p:p + nc].tolist())
        p += nc
    return out


def wfa_exact_batch_sharded(opt, pairs, align_fn=None, group=None, balance=False, device=None):
    """Align `pairs` (the same list on every rank) across all ranks of `group`; rank 0 returns the full result list
    [(s, n_cigar, n_iter, [cigar words])] in input order, the other ranks return None.

    align_fn(opt, local_pairs) -> local results; defaults to the product's mwf_wfa_exact_batch on this rank's GPU.
    """
    if align_fn is None:
        from . import api
        align_fn = api.wfa_exact_batch
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    costs = [pair_cost(t, q) for t, q in pairs] if balance else None
    idx = shard_indices(len(pairs), world, rank, costs)
    local = align_fn(opt, [pairs[i] for i in idx]) if idx else []
    return gather_to_root(idx, local, len(pairs), group=group, device=device)


def gather_to_root(idx, local, n_total, group=None, device=None, fixed=False):
    """The single end-of-batch collective: every rank contributes the records of its pairs `idx`; rank 0 returns
    the n_total results in input order, the others None.

    fixed=True (score-only batches sharded round-robin or in equal blocks: no CIGAR words, at most ceil(n_total / world)
    pairs per rank, which every rank can compute) makes it literally ONE collective -- a gather of equal-sized records;
    otherwise a tiny all_gather of the payload sizes precedes the padded gather."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return unpack_results(pack_results(idx, local), [None] * n_total)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    flat = pack_results(idx, local)
    if fixed:
        cap = 1 + 4 * ((n_total + world - 1) // world)
        assert len(flat) <= cap, "fixed-size gather: a rank holds CIGAR words or more than its share of the pairs"
        sizes = None
    else:  # sizes first (tiny all_gather), then ONE padded gather of the records to rank 0
        size = torch.tensor([len(flat)], dtype=torch.int64, device=device)
        sizes = [torch.zeros_like(size) for _ in range(world)]
        dist.all_gather(sizes, size, group=group)
        cap = max(int(s.item()) for s in sizes)
    host = torch.zeros(cap, dtype=torch.int64, pin_memory=device.type == "cuda")
    host[:len(flat)] = torch.from_numpy(flat)
    buf = host.to(device, non_blocking=True)
    gathered = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        return None
    out = [None] * n_total
    for g in range(world):
        unpack_results(gathered[g].cpu().numpy(), out)  # (the count in word 0 bounds what is read: padding is ignored)
    return out
