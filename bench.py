#!/usr/bin/env python
"""bench.py -- encode_batch throughput of the B200 engine on BASELINE.json's headline config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mb 1024] [--config gpt2|llama3|wordpiece]

One step = one pass of the whole hot path (doc_mark -> pretok_scan -> page_scan -> bpe_tile) over one batch of the
synthetic corpus of SURVEY.md §8(d) config 2 ("GPT-2 ByteLevel BPE, 1 GB synthetic UTF-8 docs avg 512 B").  Prints ONE
JSON line (rank 0).  `value` is device-resident (input already in HBM, CUDA events); `e2e` goes through the C-ABI call
b2t_encode_batch with pinned HOST buffers, copies inside the timed region; `roofline` is the pre-tokenization scan
kernel (the HBM-bound one) from CUDA events recorded on its launch stream; `cpu_baseline` is the reference's own Rust
implementation (the `tokenizers` wheel) on this box's host cores over a bounded sample.

With N > 1 (torchrun) every rank encodes its own shard of N x the corpus (weak scaling, no data-path collective: the
path shards by documents); `value` = all ranks' bytes / max-over-ranks device time.  The NCCL all-gather-v of the token
CSR that BASELINE.json's north_star mentions is timed in a second loop and reported under `allgather`;
`per_rank_ms_per_step`, `per_rank_kernel_ms_per_step` and the per-GPU clocks show where a straggler comes from.
"""
import argparse, ctypes, gzip, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402

ASSET = {"gpt2": "gpt2_style", "llama3": "llama3_style", "wordpiece": "wordpiece"}
KIND = {"gpt2": 2, "llama3": 2, "wordpiece": 4}
SEED = {"gpt2": 2, "llama3": 3, "wordpiece": 4}


def tokenizer_json(cfg):
    return gzip.open(os.path.join(ROOT, "assets", ASSET[cfg] + ".json.gz")).read().decode("utf-8")


def gen_corpus(kind, seed, first_doc, n_docs, max_bytes, out):
    """Generate docs [first_doc, ...) into `out` (np.uint8 view of pinned memory) with 8 host threads."""
    import corpus
    corpus.build()
    nthr = 8
    per = (n_docs + nthr - 1) // nthr
    parts = [None] * nthr
    cap_each = max_bytes // nthr

    def work(i):
        buf = np.empty(cap_each + 200000, dtype=np.uint8)
        d, o = corpus.generate(kind, seed, first_doc + i * per, per, max_bytes=cap_each, out=buf)
        parts[i] = (d, o)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(nthr)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    pos, offs = 0, [np.zeros(1, dtype=np.uint64)]
    for d, o in parts:
        out[pos:pos + len(d)] = d
        offs.append(o[1:] + np.uint64(pos))
        pos += len(d)
    return pos, np.concatenate(offs)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpus):
        self.gpus, self.p, self.lines = list(gpus), None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", ",".join(map(str, self.gpus)), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = {}, [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.setdefault(f[0], []).append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        med = {g: float(np.median(v)) for g, v in sorted(sm.items(), key=lambda kv: (len(kv[0]), kv[0]))}
        # sm_mhz: the slowest GPU's median under load (every rank's GPU is sampled, not only rank 0's)
        return {"sm_mhz": min(med.values()) if med else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": min((len(v) for v in sm.values()), default=0), "per_gpu_sm_mhz": list(med.values())}


def cpu_reference_worker(cfg, threads, budget_s):
    """Runs in a fresh process (rayon's pool size is fixed at first use): the reference's own Rust encode_batch
    (tokenizers wheel, bindings/python/src/tokenizer.rs:1312-1340) over a bounded sample of the bench corpus."""
    os.environ["TOKENIZERS_PARALLELISM"] = "true"
    os.environ["RAYON_NUM_THREADS"] = str(threads)
    import tokenizers
    tok = tokenizers.Tokenizer.from_str(tokenizer_json(cfg))
    cap = 96 << 20
    buf = np.empty(cap + (1 << 20), dtype=np.uint8)
    n, off = gen_corpus(KIND[cfg], SEED[cfg], 0, cap // 300, cap, buf)
    raw = buf[:n].tobytes()
    n_avail = len(off) - 1

    def docs(k):
        return [raw[int(off[i]):int(off[i + 1])].decode("utf-8") for i in range(k)]
    probe_n = min(n_avail, 16384)
    d = docs(probe_n)
    tok.encode_batch(d[:2048], add_special_tokens=False)  # warm-up (rayon pool)
    t0 = time.perf_counter(); tok.encode_batch(d, add_special_tokens=False); t1 = time.perf_counter()
    rate = int(off[probe_n]) / (t1 - t0)
    k = int(min(n_avail, max(probe_n, np.searchsorted(off, rate * budget_s))))
    d = docs(k)
    best = None
    for _ in range(2):
        t0 = time.perf_counter(); enc = tok.encode_batch(d, add_special_tokens=False); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    nbytes = int(off[k]); ntok = sum(len(e.ids) for e in enc)
    return {"value": nbytes / best / 1e9, "unit": "GB/s", "tokens_per_s": ntok / best, "cores": threads, "kind": "reference",
            "sample": f"tokenizers wheel {tokenizers.__version__} Tokenizer.encode_batch (char offsets), RAYON_NUM_THREADS={threads}, "
                      f"first {k} docs / {nbytes / 1e6:.1f} MB of the bench corpus, best of 2",
            "seconds": best}


def cpu_reference(cfg, budget_s=12.0):
    """Best of a small thread sweep (all visible cores, 32, 16, 8, 4), each in its own process."""
    cores = len(os.sched_getaffinity(0))
    tried = []
    for th in sorted({cores, min(cores, 32), min(cores, 16), min(cores, 8), min(cores, 4)}, reverse=True):
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-worker", cfg, str(th), str(budget_s)],
                                 capture_output=True, text=True, timeout=600)
            r = json.loads(out.stdout.strip().splitlines()[-1])
            tried.append(r)
        except Exception as ex:
            tried.append({"value": 0.0, "cores": th, "error": str(ex)[:200]})
    best = max(tried, key=lambda r: r.get("value") or 0.0)
    best = dict(best)
    best["host_cores"] = cores
    best["sweep"] = {str(r["cores"]): round(r.get("value") or 0.0, 5) for r in tried}
    return best


def main():
    if len(sys.argv) >= 5 and sys.argv[1] == "--cpu-worker":
        print(json.dumps(cpu_reference_worker(sys.argv[2], int(sys.argv[3]), float(sys.argv[4]))))
        return
    # Only the JSON line may reach stdout (NCCL and others print there): park the real stdout, send fd 1 to stderr.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--mb", type=int, default=1024, help="corpus size per GPU in MiB")
    ap.add_argument("--config", default="gpt2", choices=list(ASSET))
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--kind", type=int, default=0, help="corpus kind override (5 = length-skew stress of BASELINE configs[4])")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl != "reference" else a.warmup
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = a.config
    workload_kind = {5: " [length-skew corpus: Zipf doc lengths 8 B-64 KB, 0.1 % docs hold a 4-64 KB letter/space run]"}.get(a.kind, "")
    workload = {"gpt2": "GPT-2 ByteLevel BPE (50257 vocab trained offline by the reference trainer), synthetic UTF-8 docs avg ~480 B",
                "llama3": "Llama-3-style BPE (tiktoken regex, 128k vocab, ignore_merges)", "wordpiece": "Whitespace + WordPiece 30522"}[cfg]
    max_bytes = a.mb << 20
    n_docs_target = max_bytes // 300

    if a.impl == "reference":
        # the reference's CPU implementation, all host threads, bounded sample per step; rank 0 only
        if rank != 0:
            return
        per_step = []
        for s_i in range(a.warmup + a.steps):
            r = cpu_reference(cfg, budget_s=max(2.0, 45.0 / (3 * (a.warmup + a.steps))))
            if s_i >= a.warmup:
                per_step.append(r)
        r = max(per_step, key=lambda x: x["value"])
        out = {"impl": "reference", "metric": "encode_batch input throughput", "value": r["value"], "unit": "GB/s", "tokens_per_s": r["tokens_per_s"],
               "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": {"workload": workload, "bytes_per_step": None, "sample": r["sample"]},
               "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cores", "sweep")},
               "e2e": {"value": r["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        real_stdout.write(json.dumps(out) + "\n"); real_stdout.flush()
        return

    import torch
    from tokenizers_b200 import Tokenizer, _lib
    from tokenizers_b200.parallel import all_gather_csr
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    tok = Tokenizer.from_str(tokenizer_json(cfg), device=local)

    # ---- corpus: this rank's shard, generated straight into pinned host memory
    hptr = ctypes.c_void_p()
    _lib.check(L.b2t_host_alloc(max_bytes + (1 << 20), ctypes.byref(hptr)))
    hbuf = np.ctypeslib.as_array(ctypes.cast(hptr, ctypes.POINTER(ctypes.c_uint8)), shape=(max_bytes + (1 << 20),))
    kind = a.kind or KIND[cfg]
    if kind == 5:
        n_docs_target = max_bytes // 200
    n, off = gen_corpus(kind, 5 if kind == 5 else SEED[cfg], rank * n_docs_target, n_docs_target, max_bytes, hbuf)
    n_docs = len(off) - 1
    hoff_ptr = ctypes.c_void_p()
    _lib.check(L.b2t_host_alloc((n_docs + 1) * 8, ctypes.byref(hoff_ptr)))
    hoff = np.ctypeslib.as_array(ctypes.cast(hoff_ptr, ctypes.POINTER(ctypes.c_uint64)), shape=(n_docs + 1,))
    hoff[:] = off
    d_bytes = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    d_bytes[:n].copy_(torch.from_numpy(hbuf[:n]))
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    torch.cuda.synchronize()
    flags = _lib.WANT_OFFSETS
    stream = torch.cuda.current_stream()
    _lib.check(L.b2t_engine_set_profiling(tok.handle, 1))

    class DevArr:  # zero-copy torch view of an engine-owned device buffer
        def __init__(self, ptr, count, typestr):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3}

    gathered = {}

    def step_device(gather=False):
        res = ctypes.c_void_p()
        _lib.check(L.b2t_encode_batch_device(tok.handle, d_bytes.data_ptr(), n, d_off.data_ptr(), n_docs, flags, ctypes.c_void_p(stream.cuda_stream), ctypes.byref(res)))
        T = L.b2t_result_n_tokens(res)
        if gather:  # one all-gather-v of the token CSR over NCCL (tokenizers_b200/parallel.py)
            ids = torch.as_tensor(DevArr(L.b2t_result_ids(res), T, "<i4"), device="cuda")
            offs = torch.as_tensor(DevArr(L.b2t_result_offsets(res), 2 * T, "<i4"), device="cuda")
            rp = torch.as_tensor(DevArr(L.b2t_result_row_ptr(res), n_docs + 1, "<i8"), device="cuda")
            gathered["csr"] = all_gather_csr(ids, offs.reshape(-1, 2), rp)
        L.b2t_result_free(res)
        return T

    names = (ctypes.c_char_p * 16)(); ms = (ctypes.c_float * 16)()
    # one node: local rank r runs on the r-th visible GPU (nvidia-smi does not honour CUDA_VISIBLE_DEVICES, so map it)
    vis = [v.strip() for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
    sampler = ClockSampler(vis[:world] if len(vis) >= world else range(world))
    if rank == 0:
        sampler.start()  # samples every 50 ms from the warm-up through the timed device and e2e regions
    for _ in range(a.warmup):
        T = step_device()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    kern_ms = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    launches = 0
    for _ in range(a.steps):
        T = step_device()
        launches += L.b2t_engine_last_kernels(tok.handle, names, ms, 16)
        for i in range(16):
            if names[i] is None:
                break
            kern_ms.setdefault(names[i].decode(), []).append(ms[i])
        for i in range(16):
            names[i] = None
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dev_ms = ev0.elapsed_time(ev1)
    gather_ms = None
    if world > 1:
        # the same steps followed by the all-gather-v of ids / offsets / row_ptr that BASELINE.json's north_star names.
        # Reported separately (key "allgather"): the path itself has no exchange step -- every rank's slice of the CSR
        # is complete on its own -- and a replicate-everything collective necessarily grows with N.
        for _ in range(2):
            step_device(True)
        torch.cuda.synchronize(); dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(a.steps):
            step_device(True)
        g1.record(stream)
        torch.cuda.synchronize(); dist.barrier()
        gather_ms = g0.elapsed_time(g1)
        gathered.clear()

    # ---- end to end through the C ABI with pinned host buffers (H2D + kernels + D2H inside the call)
    _lib.check(L.b2t_engine_set_profiling(tok.handle, 0))

    def step_e2e():
        res = ctypes.c_void_p()
        _lib.check(L.b2t_encode_batch(tok.handle, hptr, hoff_ptr, n_docs, flags, ctypes.byref(res)))
        T = L.b2t_result_n_tokens(res)
        L.b2t_result_free(res)
        return T
    for _ in range(a.warmup):
        step_e2e()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        Te = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert Te == T
    clocks = sampler.stop() if rank == 0 else None

    # ---- reduce over ranks: time = max, work = sum
    if world > 1:
        v = torch.tensor([dev_ms, e2e_s * 1e3, gather_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        w = torch.tensor([float(n), float(T)], dtype=torch.float64, device="cuda")
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        mine = torch.tensor([dev_ms / a.steps, sum(float(np.mean(v)) for v in kern_ms.values())], dtype=torch.float64, device="cuda")
        allr = torch.empty(2 * world, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.reshape(world, 2).tolist()
        per_rank_ms = [round(x[0], 3) for x in allr]
        per_rank_kern = [round(x[1], 3) for x in allr]  # kernels only: the rest of a rank's step is host launch / sync gaps
        dev_ms, e2e_ms, gather_ms = v.tolist(); tot_bytes, tot_tok = w.tolist()
    else:
        e2e_ms, tot_bytes, tot_tok = e2e_s * 1e3, float(n), float(T)
        per_rank_ms = [round(dev_ms / a.steps, 3)]
        per_rank_kern = [round(sum(float(np.mean(v)) for v in kern_ms.values()), 3)]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_per_step = dev_ms / a.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    k1 = float(np.mean(kern_ms.get("pretok_scan", [float("nan")])))
    k1_bytes = n * 1.25 + (n / 2048) * 8  # bytes + doc_bits in, start_bits + page summaries out
    traffic = None
    try:  # dram bytes of one launch from the committed ncu --set full capture (profiles/), scaled by input size
        tj = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
        if tj.get("config") == cfg:
            traffic = tj["dram_bytes_per_input_byte"] * n
    except Exception:
        pass
    roof = {"kernel": "pretok_scan_kernel", "bound": "hbm", "achieved": k1_bytes / (k1 * 1e-3) / 1e9, "peak": peak,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback (of fallback)",
            "unit": "GB/s", "frac": k1_bytes / (k1 * 1e-3) / 1e9 / peak, "traffic": traffic, "algorithmic_bytes_per_launch": k1_bytes, "ms_per_launch": k1,
            "input_GBps": n / (k1 * 1e-3) / 1e9,
            "note": "achieved uses THIS kernel's layout (bytes + doc bitmap in, split bitmap + page summaries out = 1.25 B per input byte); "
                    "with SURVEY.md 8(d)'s u32-start-list accounting (N + 4*N_pretok ~ 1.75 B/B) the same time would read frac x 1.4"}
    out = {"metric": "encode_batch input throughput", "value": tot_bytes / (ms_per_step * 1e-3) / 1e9, "unit": "GB/s",
           "tokens_per_s": tot_tok / (ms_per_step * 1e-3), "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": {"workload": workload + workload_kind + f", {n / 1e6:.0f} MB / {n_docs} docs per GPU, ids + char offsets", "bytes_per_gpu": n, "docs_per_gpu": n_docs,
                      "tokens_per_gpu": int(T), "l2": "inputs larger than L2 (no flush needed)",
                      "parallelism": f"docs sharded over {world} rank(s), no data-path collective" + (" (all-gather-v variant under 'allgather')" if world > 1 else "")},
           "kernels_ms": {k: float(np.mean(v)) for k, v in kern_ms.items()}, "per_rank_ms_per_step": per_rank_ms,
           "per_rank_kernel_ms_per_step": per_rank_kern,
           "roofline": roof,
           "e2e": {"value": tot_bytes / (e2e_ms / a.steps * 1e-3) / 1e9, "unit": "GB/s", "tokens_per_s": tot_tok / (e2e_ms / a.steps * 1e-3), "ms_per_step": e2e_ms / a.steps,
                   "h2d_bytes_per_step": int(n + (n_docs + 1) * 8), "d2h_bytes_per_step": int(T * 12 + (n_docs + 1) * 8 + 16 * ((n >> 26) + 1))},
           "gpu_launches": int(launches), "clocks": clocks}
    if gather_ms is not None:
        out["allgather"] = {"what": "same steps + NCCL all-gather-v of ids/offsets/row_ptr to every rank (tokenizers_b200/parallel.py)",
                            "ms_per_step": gather_ms / a.steps, "value": tot_bytes / (gather_ms / a.steps * 1e-3) / 1e9, "unit": "GB/s"}
    if not a.no_cpu and world == 1:
        try:
            out["cpu_baseline"] = {k: v for k, v in cpu_reference(cfg).items() if k != "seconds"}
        except Exception as ex:  # the wheel is part of the image; if it is missing say so instead of inventing a number
            out["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": len(os.sched_getaffinity(0)), "kind": "reference", "sample": f"unavailable: {ex}"}
    real_stdout.write(json.dumps(out) + "\n"); real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
