#!/usr/bin/env python
"""Developer tool: A/B several builds of libb2t.so in ONE process on the same corpus (no torch import).

    python tools/ab_kernels.py [--mb 512] [--config gpt2] [--runs 4] lib1.so lib2.so ...

Every library encodes the same batch through the C ABI (b2t_encode_batch, one chunk, profiling on); prints the best
per-kernel CUDA-event times and a checksum of ids / offsets / row_ptr, which must be the same for all libraries.
"""
import argparse, ctypes, hashlib, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=512)
    ap.add_argument("--config", default="gpt2")
    ap.add_argument("--runs", type=int, default=4)
    ap.add_argument("libs", nargs="+")
    a = ap.parse_args()
    os.environ["B2T_CHUNK_BYTES"] = str((1 << 31) - 4096)  # one chunk: the kernel events then cover the whole batch
    import bench
    from tokenizers_b200 import _lib, tokenizer
    cfg = a.config
    max_bytes = a.mb << 20
    buf = np.empty(max_bytes + (1 << 20), dtype=np.uint8)
    n, off = bench.gen_corpus(bench.KIND[cfg], bench.SEED[cfg], 0, max_bytes // 300, max_bytes, buf)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n_docs = len(off) - 1
    js = bench.tokenizer_json(cfg)
    out = {}
    for path in a.libs:
        _lib._lib = None
        _lib.LIB_PATH = os.path.abspath(path)
        L = _lib.lib()
        tok = tokenizer.Tokenizer.from_str(js)
        _lib.check(L.b2t_engine_set_profiling(tok.handle, 1))
        names = (ctypes.c_char_p * 16)(); ms = (ctypes.c_float * 16)()
        best, digest = {}, None
        for r in range(a.runs):
            res = ctypes.c_void_p()
            _lib.check(L.b2t_encode_batch(tok.handle, buf.ctypes.data, off.ctypes.data, n_docs, _lib.WANT_OFFSETS, ctypes.byref(res)))
            for i in range(16):
                names[i] = None
            L.b2t_engine_last_kernels(tok.handle, names, ms, 16)
            for i in range(16):
                if names[i] is None:
                    break
                k = names[i].decode()
                best[k] = min(best.get(k, 1e30), ms[i])
            if r == 0:
                T = L.b2t_result_n_tokens(res)
                h = hashlib.sha256()
                for ptr, cnt, dt in ((L.b2t_result_ids(res), T, np.uint32), (L.b2t_result_offsets(res), 2 * T, np.uint32),
                                     (L.b2t_result_row_ptr(res), n_docs + 1, np.uint64)):
                    arr = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(cnt,))
                    h.update(arr.tobytes())
                digest = h.hexdigest()[:16]
            L.b2t_result_free(res)
        del tok
        out[os.path.basename(path)] = {"kernels_ms": {k: round(v, 4) for k, v in best.items()}, "sum_ms": round(sum(best.values()), 3),
                                       "n_tokens": int(T), "sha": digest}
        print(os.path.basename(path), json.dumps(out[os.path.basename(path)]), flush=True)
    shas = {v["sha"] for v in out.values()}
    print("outputs identical across libraries:", len(shas) == 1)


if __name__ == "__main__":
    main()
