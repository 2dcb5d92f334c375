"""Multi-GPU plumbing for the encode_batch path: one process per GPU (torch.distributed), documents sharded by
contiguous ranges, one all-gather-v of the token CSR at the end (BASELINE.json north_star).

The path has no exchange step other than that final gather: documents are independent units (the reference itself only
parallelises over batch items, tokenizer/mod.rs:1345-1348).

`encode_batch_sharded` is the product entry point (one call per rank, every rank ends with the CSR of the whole batch):
the engine stops at the token counts (b2t_encode_batch_device_begin), the ranks exchange them, every rank's compaction
kernel then writes its tokens directly at its displacement of the gathered buffers (b2t_encode_batch_device_finish) and one
group of NCCL send / recv pairs completes the other ranks' parts in place -- no padding, no concatenation, one small host
sync for the counts.  `all_gather_csr` (padded all_gather_into_tensor + torch.cat) is the backend-neutral form the gloo
tests use.
"""
import ctypes
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


def bind_to_gpu_numa_node(device):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs), BEFORE any pinned allocation: pinned
    pages then land on that node and host<->device copies do not cross the socket link.  Returns the node or None."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device).pci_bus_id if hasattr(torch.cuda.get_device_properties(device), "pci_bus_id") else None
        dom = getattr(torch.cuda.get_device_properties(device), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device), "pci_device_id", 0)
        if bus is None:
            return None
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = set(os.sched_getaffinity(0)) & set(cpus)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


class ShardedResult:
    """The token CSR of the whole batch on this rank's GPU: ids [T], offsets [T, 2] or None, row_ptr [n_docs_total + 1]."""
    def __init__(self, ids, offsets, row_ptr, counts, doc_counts):
        self.ids, self.offsets, self.row_ptr, self.token_counts, self.doc_counts = ids, offsets, row_ptr, counts, doc_counts


def encode_batch_sharded(tok, d_bytes, n_bytes, d_doc_off, n_docs, want_offsets=True, group=None, out=None, stream=None, gather_stream=None):
    """Encode this rank's shard (device tensors: packed bytes, shard-relative doc offsets) and return the CSR of the WHOLE
    batch (all ranks' shards in rank order) on every rank.

    out: optional (ids, offsets, row_ptr) tensors to reuse (capacity checked).  gather_stream: a torch.cuda.Stream for the
    exchange -- the send / recv group then overlaps whatever the caller launches next on the current stream; the returned
    tensors are ready once that stream is."""
    from . import _lib
    L = _lib.lib()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = d_bytes.device
    st = stream or torch.cuda.current_stream(dev)
    flags = _lib.WANT_OFFSETS if want_offsets else 0
    nt = ctypes.c_uint64(0)
    _lib.check(L.b2t_encode_batch_device_begin(tok.handle, d_bytes.data_ptr(), n_bytes, d_doc_off.data_ptr(), n_docs, flags,
                                               ctypes.c_void_p(st.cuda_stream), ctypes.byref(nt)))
    mine = torch.tensor([nt.value, n_docs], dtype=torch.int64, device=dev)
    allc = torch.empty(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allc, mine, group=group)
    allc = allc.reshape(world, 2).tolist()          # the one host sync of the call
    tcnt = [int(x[0]) for x in allc]; dcnt = [int(x[1]) for x in allc]
    tdisp = [0]; ddisp = [0]
    for r in range(world):
        tdisp.append(tdisp[-1] + tcnt[r]); ddisp.append(ddisp[-1] + dcnt[r])
    T, D = tdisp[-1], ddisp[-1]
    if out is not None and out[0].numel() >= T and out[2].numel() >= D + 1 and (not want_offsets or out[1].numel() >= 2 * T):
        ids, offs, rp = out
    else:
        ids = torch.empty(T, dtype=torch.int32, device=dev)
        offs = torch.empty((T, 2), dtype=torch.int32, device=dev) if want_offsets else None
        rp = torch.empty(D + 1, dtype=torch.int64, device=dev)
    offs_flat = offs.reshape(-1) if want_offsets else None
    if gather_stream is not None and gather_stream is not st:
        st.wait_stream(gather_stream)   # an exchange of an earlier call may still be using these buffers
    # my part, written in place by the compaction kernel (row_ptr[ddisp .. ddisp + n_docs] includes the next rank's first entry)
    _lib.check(L.b2t_encode_batch_device_finish(
        tok.handle, ids.data_ptr() + 4 * tdisp[rank], (offs_flat.data_ptr() + 8 * tdisp[rank]) if want_offsets else None, None,
        rp.data_ptr() + 8 * ddisp[rank], tdisp[rank], ctypes.c_void_p(st.cuda_stream)))
    if world > 1:
        gs = gather_stream or st
        if gs is not st:
            gs.wait_stream(st)
        with torch.cuda.stream(gs):
            ops = []
            for peer in range(world):
                if peer == rank:
                    continue
                if tcnt[rank]:
                    ops.append(dist.P2POp(dist.isend, ids[tdisp[rank]:tdisp[rank + 1]], peer, group))
                    if want_offsets:
                        ops.append(dist.P2POp(dist.isend, offs_flat[2 * tdisp[rank]:2 * tdisp[rank + 1]], peer, group))
                if dcnt[rank]:
                    ops.append(dist.P2POp(dist.isend, rp[ddisp[rank] + 1:ddisp[rank + 1] + 1], peer, group))
                if tcnt[peer]:
                    ops.append(dist.P2POp(dist.irecv, ids[tdisp[peer]:tdisp[peer + 1]], peer, group))
                    if want_offsets:
                        ops.append(dist.P2POp(dist.irecv, offs_flat[2 * tdisp[peer]:2 * tdisp[peer + 1]], peer, group))
                if dcnt[peer]:
                    ops.append(dist.P2POp(dist.irecv, rp[ddisp[peer] + 1:ddisp[peer + 1] + 1], peer, group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            if rank != 0:
                rp[0:1].zero_()
    return ShardedResult(ids[:T], offs[:T] if want_offsets else None, rp[:D + 1], tcnt, dcnt)
