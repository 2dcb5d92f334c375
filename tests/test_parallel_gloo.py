"""CPU: the N>1 plumbing (document sharding + all-gather-v of the token CSR) on the gloo backend, world_size 2.
The shards are encoded with the ORACLE here (no GPU in this test); what is checked is that gathering the per-rank CSRs
reproduces the single-process CSR of the whole batch, i.e. N-rank result == 1-rank result."""
import os, sys, socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import helpers, corpus


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tokenizers_b200.parallel import shard_range, all_gather_csr
        from oracle import oracle as orc
        o = orc.Oracle(helpers.asset_json("gpt2_style"))
        data, off = corpus.generate(2, 123, 0, 301)
        n = len(off) - 1
        lo, hi = shard_range(n, rank, world)
        sl = data[int(off[lo]):int(off[hi])]
        ids, offs, wid, rp = o.encode_batch_csr(sl, (off[lo:hi + 1] - off[lo]).astype(np.uint64))
        g_ids, g_off, g_rp = all_gather_csr(torch.from_numpy(ids.astype(np.int64)), torch.from_numpy(offs.astype(np.int64)),
                                            torch.from_numpy(rp.astype(np.int64)))
        if rank == 0:
            e_ids, e_offs, _, e_rp = o.encode_batch_csr(data, off)
            ret["ok"] = bool(np.array_equal(g_ids.numpy(), e_ids) and np.array_equal(g_off.numpy(), e_offs) and
                             np.array_equal(g_rp.numpy(), e_rp.astype(np.int64)))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    from tokenizers_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_shard_by_bytes_balances_skewed_lengths():
    from tokenizers_b200.parallel import shard_by_bytes
    rng = np.random.default_rng(5)
    lens = np.minimum((rng.pareto(1.2, 20000) * 50 + 8).astype(np.int64), 65536)  # Zipf-like lengths, 64 KB outliers
    off = np.concatenate([[0], np.cumsum(lens)])
    for w in (1, 2, 3, 8):
        r = shard_by_bytes(off, w)
        assert r[0][0] == 0 and r[-1][1] == len(lens) and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        by = [int(off[b] - off[a]) for a, b in r]
        assert max(by) - min(by) <= 2 * int(lens.max()), by
    assert shard_by_bytes(np.array([0]), 4) == [(0, 0)] * 4 and shard_by_bytes(np.array([0, 5]), 2) in ([(0, 0), (0, 1)], [(0, 1), (1, 1)])


def test_two_rank_gather_equals_single_rank():
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret.get("ok") is True
