"""Multi-GPU plumbing for the encode_batch path: one process per GPU (torch.distributed), documents sharded by
contiguous ranges, one all-gather-v of the token CSR at the end (BASELINE.json north_star).

The path has no exchange step other than that final gather: documents are independent units (the reference itself only
parallelises over batch items, tokenizer/mod.rs:1345-1348).  The helpers work on any backend (NCCL on GPUs, gloo in the
CPU tests) because they only use all_gather_into_tensor / all_gather on padded tensors.
"""
import torch
import torch.distributed as dist


def shard_range(n_docs, rank, world):
    """Contiguous, balanced document range [lo, hi) of `rank` (first n_docs % world ranks get one extra)."""
    base, rem = divmod(n_docs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_bytes(doc_off, world):
    """Contiguous document ranges with (nearly) equal BYTES per rank: [(lo, hi)] * world.

    Equal document counts are only balanced when documents have similar lengths; with the length-skewed corpus of
    BASELINE config 5 (Zipf lengths, 64 KB outliers) the per-rank work follows the bytes.  Cut points
    are the document boundaries closest to the multiples of total / world; ranges stay contiguous, so the gathered CSR
    is in input order without a permutation."""
    import numpy as np
    off = np.asarray(doc_off, dtype=np.int64)
    n, total = len(off) - 1, int(off[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        k = int(np.searchsorted(off, target, side="left"))
        if k > 0 and k <= n and abs(int(off[k - 1]) - target) <= abs(int(off[min(k, n)]) - target):
            k -= 1
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def all_gather_v(local, group=None):
    """All-gather 1-D tensors of different lengths.  Returns (flat tensor of every rank's data in rank order, counts)."""
    world = dist.get_world_size(group)
    cnt = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    cnts = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    counts = cnts.tolist()
    mx = max(counts) if counts else 0
    padded = local
    if local.numel() < mx:
        padded = torch.zeros(mx, dtype=local.dtype, device=local.device)
        padded[: local.numel()] = local
    out = torch.empty(world * mx, dtype=local.dtype, device=local.device)
    if mx:
        dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    flat = torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)]) if mx else out
    return flat, counts


def all_gather_csr(ids, offsets, row_ptr, group=None):
    """Gather the per-rank CSRs (ids [T], offsets [T, 2] or None, row_ptr [n+1] local) into the global CSR, rank order.

    row_ptr is rebased by the exclusive scan of the ranks' token totals, so that the result equals the CSR a single rank
    would have produced for the concatenated batch."""
    g_ids, counts = all_gather_v(ids, group)
    g_off = None
    if offsets is not None:
        flat, _ = all_gather_v(offsets.reshape(-1), group)
        g_off = flat.reshape(-1, 2)
    rp_local = row_ptr[1:].to(torch.int64)  # drop each rank's leading 0
    g_rp, rp_counts = all_gather_v(rp_local, group)
    out, base, pos = [torch.zeros(1, dtype=torch.int64, device=ids.device)], 0, 0
    for r, c in enumerate(rp_counts):
        out.append(g_rp[pos: pos + c] + base)
        base += counts[r]
        pos += c
    return g_ids, g_off, torch.cat(out)
