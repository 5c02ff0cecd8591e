"""SASS opcode mix of one kernel from `ncu -i x.ncu-rep --page source --csv` (tools/ncu_capture.sh writes that CSV):
warp instructions executed and stall samples per opcode, the FP64 share, the executed FP64 FLOPs (DFMA = 2 per lane,
DMUL / DADD = 1), and the stall-reason totals.   usage: python tools/sass_mix.py <source.csv> [top_n]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    rows = list(csv.reader(open(path, newline="")))
    kernel = rows[0][1] if rows and rows[0] and rows[0][0] == "Kernel Name" else "?"
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    inst, samp, thr = defaultdict(float), defaultdict(float), defaultdict(float)
    stalls = defaultdict(float)
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    n_sass = 0
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        src = r[col["Source"]].strip()
        if not src:
            continue
        tok = src.split()
        op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
        op = op.rstrip(";")
        base = op.split(".")[0]
        if base == "MUFU":
            base = op
        n_sass += 1
        inst[base] += float(r[col["Instructions Executed"]] or 0)
        thr[base] += float(r[col["Thread Instructions Executed"]] or 0)
        samp[base] += float(r[col["# Samples"]] or 0)
        for s in stall_cols:
            stalls[s] += float(r[col[s]] or 0)
    ti, ts = sum(inst.values()), sum(samp.values())
    print("kernel: %s" % kernel)
    print("%-16s %16s %7s %12s %7s" % ("opcode", "warp instr", "share", "samples", "share"))
    for op, n in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
        print("%-16s %16d %6.1f%% %12d %6.1f%%" % (op, n, 100 * n / ti, samp[op], 100 * samp[op] / max(ts, 1)))
    fp64 = sum(inst[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    flops = 2 * thr["DFMA"] + thr["DMUL"] + thr["DADD"]
    print("total %d warp instructions in %d SASS instructions; FP64 arithmetic (DFMA+DMUL+DADD+DSETP+DMNMX) %.1f%%; "
          "DFMA alone %.1f%%; DFMA share of the FP64 issue slots %.1f%%; MUFU %.2f%%" % (
              ti, n_sass, 100 * fp64 / ti, 100 * inst["DFMA"] / ti, 100 * inst["DFMA"] / max(fp64, 1),
              100 * sum(v for k, v in inst.items() if k.startswith("MUFU")) / ti))
    print("executed FP64 FLOPs (thread-level: 2 DFMA + DMUL + DADD): %.4e" % flops)
    tot = sum(stalls.values())
    print("stall samples: " + ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(tot, 1)) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]))


if __name__ == "__main__":
    main()
