#!/usr/bin/env python
"""Turn the files `tools/gpu_round.sh <tag>` left in gpurun_out/ into the tracked summaries under profiles/:

    python tools/refresh_profiles.py r02

  profiles/bench_<tag>_final.json (+ _llama3 / _wordpiece)   the bench lines
  profiles/launches_<tag>.csv, launch_shares_<tag>.txt       ncu launch list of the bench command + kernel shares
  profiles/prof_k1_<tag>_summary.txt, prof_k2_<tag>_summary.txt   metrics + stall reasons + hottest source lines
  profiles/k1_traffic.json                                   DRAM bytes per input byte of the scan kernel (bench.py reads it)
Needs ncu / cuobjdump / nvdisasm on PATH (they are in the dev image) and the libb2t.so the captures were taken with.
"""
import collections, csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, capture_output=True, text=True, **kw).stdout


def launch_shares(tag):
    src = os.path.join(OUT, f"{tag}_launches.csv")
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    by = collections.defaultdict(list)
    for r in rows:
        v = float(r[-1].replace(",", "")) / {"ns": 1e6, "us": 1e3, "usecond": 1e3, "msecond": 1.0, "ms": 1.0}.get(r[-2], 1e6)
        by[re.sub(r"\(.*", "", r[4]).replace("b2t::", "")].append((r[8], v))
    out = []
    for k, v in by.items():  # full-batch launches = the largest grid of each kernel (the e2e leg launches 64 MB chunks too)
        g = max({x[0] for x in v}, key=lambda gs: int(gs.strip("()").split(",")[0]))
        sel = [x[1] for x in v if x[0] == g]
        out.append((sum(sel) / len(sel), len(sel), k))
    tot = sum(o[0] for o in out)
    lines = [f"{100 * o[0] / tot:5.1f}%  {o[0]:9.3f} ms avg  x{o[1]}  {o[2]}" for o in sorted(out, reverse=True)]
    hdr = ("ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 3 --no-cpu;\n"
           "full-batch launches only; serialised, cold-cache times\n")
    open(os.path.join(PROF, f"launch_shares_{tag}.txt"), "w").write(hdr + "\n".join(lines) + "\n")
    shutil.copy(src, os.path.join(PROF, f"launches_{tag}.csv"))
    print("\n".join(lines[:4]))


KEYS = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]


def kernel_summary(tag, which, kname, sass):
    rep = os.path.join(OUT, f"{tag}_{which}.ncu-rep")
    if not os.path.exists(rep):
        print("missing", rep); return None
    raw = list(csv.reader(run(["ncu", "-i", rep, "--page", "raw", "--csv"]).splitlines()))
    h, u, r = raw[0], raw[1], raw[2]
    d, un = dict(zip(h, r)), dict(zip(h, u))
    lines = [f"# ncu --set full --clock-control none --import-source on, {which}, bench.py --mb 256, {tag}"]
    lines += [f"{k} = {d[k]} {un.get(k, '')}" for k in KEYS if k in d]
    for k in h:
        if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and d[k] and float(d[k]) > 0.5:
            lines.append("stall " + k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "") + f" {float(d[k]):.6f}")
    src_csv = os.path.join(OUT, f"{tag}_{which}_src.csv")
    open(src_csv, "w").write(run(["ncu", "-i", rep, "--page", "source", "--csv"]))
    lines.append(run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), src_csv, kname, sass, "25"]))
    open(os.path.join(PROF, f"prof_{which}_{tag}_summary.txt"), "w").write("\n".join(lines))
    print("\n".join(lines[1:6]))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    d["_dram_bytes"] = sum(float(d[k]) * scale.get(un.get(k, "byte"), 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    return d


def main():
    tag = sys.argv[1]
    for suffix in ("", "_llama3", "_wordpiece"):
        src = os.path.join(OUT, f"{tag}_bench{suffix}.json")
        if os.path.exists(src) and os.path.getsize(src):
            shutil.copy(src, os.path.join(PROF, f"bench_{tag}{suffix or '_final'}.json"))
    launch_shares(tag)
    work = os.path.join(ROOT, "build", "sass_" + tag)
    os.makedirs(work, exist_ok=True)
    run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "tokenizers_b200", "libb2t.so")], cwd=work)
    sass = os.path.join(work, "engine.sass")
    open(sass, "w").write(run(["nvdisasm", "-g", os.path.join(work, "engine.sm_100a.cubin")]))
    d = kernel_summary(tag, "k1", "pretok_lean_kernelILi0", sass)
    kernel_summary(tag, "k2", "model_tile_kernelILi0", sass)
    if d:
        n_in = 256 * 2 ** 20  # bench.py --mb 256 (the generator stops a little short of it; good to three digits)
        per = d["_dram_bytes"] / n_in
        json.dump({"config": "gpt2", "dram_bytes_per_input_byte": round(per, 4),
                   "source": f"ncu --set full, pretok_lean_kernel<0>, bench.py --mb 256 (gpurun_out/{tag}_k1.ncu-rep)"},
                  open(os.path.join(PROF, "k1_traffic.json"), "w"))


if __name__ == "__main__":
    main()
