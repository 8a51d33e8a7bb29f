#!/usr/bin/env python
"""Per-source-line warp-stall samples of one kernel from an .ncu-rep (captured with
--import-source on / -lineinfo).  ncu's CSV source page is SASS-only, so the SASS offsets are
joined with `nvdisasm -gi` line info of the cubin inside the shipped .so.

    python tools/ncu_hotspots.py REPORT.ncu-rep KERNEL_MANGLED_SUBSTR [--top 40] [--outer FILE]

--outer FILE attributes inlined code to the line of FILE at the bottom of its inline chain
(phase-level view); default is the innermost line."""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    table, cur, on, after_ins = {}, [], False, True
    re_file = re.compile(r'//## File "([^"]+)", line (\d+)')
    re_ins = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);")
    for ln in txt.splitlines():
        if ln.startswith("\t.section") or ln.startswith(".section"):
            on = (".text." in ln) and (kernel in ln)
            continue
        if not on:
            continue
        m = re_file.search(ln)
        if m:   # consecutive File lines = inline chain, innermost first, outermost last
            if after_ins:
                cur = []
                after_ins = False
            ent = (os.path.basename(m.group(1)), int(m.group(2)))
            if not cur or cur[-1] != ent:
                cur.append(ent)
            continue
        m = re_ins.search(ln)
        if m:
            after_ins = True
            table[int(m.group(1), 16)] = (list(cur), m.group(2).split()[0])
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--so", default=os.path.join(ROOT, "scale_letkf_b200", "libletkf_b200.so"))
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--outer", default=None)
    args = ap.parse_args()
    table = line_table(args.so, args.kernel)
    out = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci = {n: hdr.index(n) for n in hdr}
    base = None
    agg = collections.Counter()
    inst = collections.Counter()
    stall = collections.defaultdict(collections.Counter)
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    total = 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[0], 16)
        if base is None:
            base = addr
        chain, _op = table.get(addr - base, ([("?", 0)], "?"))
        if not chain:
            chain = [("?", 0)]
        key = chain[0]
        if args.outer:
            outer = [c for c in chain if c[0] == args.outer]
            key = outer[-1] if outer else chain[-1]
        n = int(r[ci["# Samples"]] or 0)
        agg[key] += n
        inst[key] += int(r[ci["Instructions Executed"]] or 0)
        total += n
        for s in stall_cols:
            v = int(r[ci[s]] or 0)
            if v:
                stall[key][s[6:]] += v
    print(f"total samples {total}")
    print("| file:line | samples | share | warp-instr | top stalls |")
    print("|---|---|---|---|---|")
    for key, n in agg.most_common(args.top):
        st = ", ".join(f"{k} {v}" for k, v in stall[key].most_common(3))
        print(f"| {key[0]}:{key[1]} | {n} | {100.0 * n / max(total, 1):.1f}% | {inst[key]} | {st} |")


if __name__ == "__main__":
    main()
