#!/usr/bin/env python
"""Predict the dynamic warp-instruction count of a rebuilt kernel without a GPU.

    sass_predict.py <ncu source csv> <profiled sass> <profiled csrc dir> <new sass> <new csrc dir> <kernel substring>

The ncu capture of the PROFILED build gives, per source line, how often its SASS instructions executed (executed
warp-instructions / static instructions of that line).  Lines are matched between the two builds by their TEXT, so
the prediction is  sum over lines of  static_instructions_new(line) x frequency_profiled(line text);  a line the
profiled build does not have inherits the frequency of the closest earlier known line of the same file.  It is a model
(the compiler may move work between lines), good for ranking edits before spending GPU time -- the number that counts
is still the CUDA-event time on the GPU.  sass files: `nvdisasm -g <cubin>` of `cuobjdump -xelf all libb2t.so`.
"""
import csv, re, sys, os


def parse_sass(path, kname):
    """-> list of (file, line, opcode) for the instructions of the first function whose name contains kname"""
    out, infn, cur = [], False, None
    for ln in open(path):
        if ln.startswith("//-----"):
            infn = kname in ln and ".text." in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
        if m:
            out.append((cur[0] if cur else None, cur[1] if cur else 0, m.group(1)))
    return out


def load_src(d):
    cache = {}

    def text(f, l):
        if f not in cache:
            try:
                cache[f] = open(os.path.join(d, f)).read().split("\n")
            except OSError:
                cache[f] = []
        return cache[f][l - 1].strip() if 0 < l <= len(cache[f]) else ""
    return text


ALU = ("LOP3", "SHF", "IADD3", "PRMT", "ISETP", "SEL", "IMNMX", "VIMNMX", "MOV", "LEA", "BMSK", "SGXT", "IABS", "FMNMX", "PLOP3", "VABSDIFF", "R2P", "P2R")
FMA = ("IMAD", "FFMA", "FMUL", "HFMA2", "IMUL", "IDP")
XU = ("POPC", "FLO", "BREV", "MUFU", "I2F", "F2I")


def pipe(op):
    base = op.split(".")[0]
    if base in ALU: return "alu"
    if base in FMA: return "fma"
    if base in XU: return "xu"
    if base.startswith(("LD", "ST", "ATOM", "RED", "SHFL", "VOTE", "MATCH", "REDUX")): return "mem/warp"
    return "other"


def main():
    csv_path, sass_old, dir_old, sass_new, dir_new, kname = sys.argv[1:7]
    rows = list(csv.reader(open(csv_path)))
    h = rows[1]
    ix = h.index("Instructions Executed")
    data = [int(r[ix]) for r in rows[2:] if len(r) > ix and r[ix].isdigit()]
    old = parse_sass(sass_old, kname)
    new = parse_sass(sass_new, kname)
    if len(old) != len(data):
        sys.exit(f"profiled sass has {len(old)} instructions but the ncu csv has {len(data)} rows: not the same build")
    t_old, t_new = load_src(dir_old), load_src(dir_new)
    freq, stat_old = {}, {}
    for (f, l, op), ex in zip(old, data):
        k = (f, t_old(f, l))
        a = freq.setdefault(k, [0, 0]); a[0] += ex; a[1] += 1
    total_old = sum(data)
    pred, by_pipe, unknown = 0.0, {}, 0
    last = {}
    per_line = {}
    for f, l, op in new:
        k = (f, t_new(f, l))
        if k in freq:
            fr = freq[k][0] / freq[k][1]; last[f] = fr
        else:
            fr = last.get(f, 0.0); unknown += 1
        pred += fr
        by_pipe[pipe(op)] = by_pipe.get(pipe(op), 0.0) + fr
        per_line[k] = per_line.get(k, 0.0) + fr
    old_pipe = {}
    for (f, l, op), ex in zip(old, data):
        old_pipe[pipe(op)] = old_pipe.get(pipe(op), 0) + ex
    print(f"profiled: {len(old)} static instr, {total_old} executed")
    print(f"new:      {len(new)} static instr, predicted {pred:.0f} executed ({100.0 * pred / total_old:.1f} % of profiled), "
          f"{unknown} instr on lines the profile does not know")
    for p in sorted(set(by_pipe) | set(old_pipe)):
        print(f"  pipe {p:9s} profiled {100.0 * old_pipe.get(p, 0) / total_old:5.1f} %   new {100.0 * by_pipe.get(p, 0) / total_old:5.1f} % (of profiled total)")
    if len(sys.argv) > 7:
        print("largest per-line changes (new - profiled, % of profiled total):")
        old_line = {k: v[0] for k, v in freq.items()}
        keys = set(old_line) | set(per_line)
        diffs = sorted(((per_line.get(k, 0.0) - old_line.get(k, 0)) / total_old * 100.0, k) for k in keys)
        for d, k in diffs[:int(sys.argv[7])] + diffs[-int(sys.argv[7]):]:
            print(f"  {d:+6.2f}  {k[0]}: {k[1][:100]}")


if __name__ == "__main__":
    main()
