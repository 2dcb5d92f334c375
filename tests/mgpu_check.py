"""Run under torchrun (one process per GPU): the CSR that N ranks produce together (encode_batch_sharded: byte-balanced
contiguous shards, compaction straight into the gathered buffers, one send/recv group) must equal the CSR one GPU
produces for the whole batch.  Used by tests/test_gpu_parity.py::test_multi_gpu_sharded_equals_single when >= 2 GPUs
are visible, and by hand:  torchrun --standalone --nproc-per-node 2 tests/mgpu_check.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
import helpers, corpus  # noqa: E402
from tokenizers_b200 import Tokenizer  # noqa: E402
from tokenizers_b200.parallel import shard_by_bytes, encode_batch_sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for name, kind in (("gpt2_style", 5), ("llama3_style", 2), ("wordpiece", 4)):
        tok = Tokenizer.from_str(helpers.asset_json(name), device=local)
        data, off = corpus.generate(kind, 77, 0, 6000)          # every rank generates the same batch
        off = off.astype(np.uint64)
        full = tok.encode_batch_csr(data, off)                  # one GPU, whole batch
        lo, hi = shard_by_bytes(off, world)[rank]
        b0, b1 = int(off[lo]), int(off[hi])
        d_bytes = torch.zeros(b1 - b0 + 64, dtype=torch.uint8, device="cuda")
        d_bytes[: b1 - b0].copy_(torch.from_numpy(data[b0:b1].copy()))
        d_off = torch.from_numpy((off[lo:hi + 1] - off[lo]).astype(np.int64)).cuda()
        for rep in range(2):                                    # second call reuses the buffers
            res = encode_batch_sharded(tok, d_bytes, b1 - b0, d_off, hi - lo, want_offsets=True, out=None if rep == 0 else (res.ids, res.offsets, res.row_ptr))
            torch.cuda.synchronize()
            ok = (np.array_equal(res.ids.cpu().numpy().view(np.uint32), full.ids) and
                  np.array_equal(res.offsets.cpu().numpy().view(np.uint32), full.offsets) and
                  np.array_equal(res.row_ptr.cpu().numpy().astype(np.uint64), full.row_ptr.astype(np.uint64)))
            assert ok, f"rank {rank}: sharded CSR of {name} differs from the single-GPU CSR"
        if rank == 0:
            print(f"{name}: {world} ranks == 1 rank ({len(full.ids)} tokens, {len(off) - 1} docs, shard token counts {res.token_counts})")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
