"""Attribute the warp-stall samples and executed instructions of an ncu capture of `k_hmc_step<16>`
to CUDA source lines (innermost inlined location), offline:

    python profiles/hot_lines.py gpurun_out/hmc_step_r1h.ncu-rep > profiles/r01_hot_lines.md

Joins ncu's SASS page (`ncu -i REP --page source --csv`: address, samples, instructions executed)
with the line table of the in-tree library (`cuobjdump -xelf` + `nvdisasm -g`), which must be the
build that was profiled.  No GPU needed.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fab_torch_b200", "csrc", "libfab_b200.so")
KERNEL = os.environ.get("FAB_HOT_KERNEL", "_Z10k_hmc_stepILi16E")     # e.g. _Z12k_hmc_step_uILi5E


def line_table():
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=d, check=True, capture_output=True)
        txt = []
        for cubin in sorted(f for f in os.listdir(d) if f.endswith(".cubin")):   # one per translation unit
            t = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True,
                               text=True).stdout
            if ".text." + KERNEL in t:
                txt = t.split("\n")
                break
    table, cur, inside = {}, ("?", 0), False
    for l in txt:
        if l.startswith(".text." + KERNEL):
            inside = True
            continue
        if inside and (l.startswith("//-----") or l.startswith(".text._Z")):
            break
        if not inside:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > ci["# Samples"]]
    base = int(data[0][ci["Address"]], 16)
    table = line_table()
    by_line = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
    by_file = collections.defaultdict(lambda: [0.0, 0.0])
    tot_s = tot_i = 0.0
    missing = 0
    for r in data:
        off = int(r[ci["Address"]], 16) - base
        s, n = float(r[ci["# Samples"]] or 0), float(r[ci["Instructions Executed"]] or 0)
        loc, sass = table.get(off, (("?", 0), ""))
        if off not in table:
            missing += 1
        op = (sass.split()[1] if sass.startswith("@") else (sass.split()[0] if sass else "?")).split(".")[0]
        e = by_line[loc]
        e[0] += s; e[1] += n; e[2][op] += n
        by_file[loc[0]][0] += s; by_file[loc[0]][1] += n
        tot_s += s; tot_i += n
    src_cache = {}

    def src(loc):
        f, ln = loc
        for d in ("fab_torch_b200/csrc", "include"):
            p = os.path.join(ROOT, d, f)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().split("\n")
                return src_cache[p][ln - 1].strip()[:90] if 0 < ln <= len(src_cache[p]) else ""
        return ""

    print(f"# {KERNEL}: warp-stall samples and executed instructions by source line\n")
    print(f"capture `{os.path.basename(rep)}`; {int(tot_s)} samples, {tot_i / 1e6:.1f} M warp-instructions; "
          f"{missing} SASS rows without a line-table entry.  Innermost (inlined) location per instruction.\n")
    print("## By file\n\n| file | samples | instructions |\n|---|---|---|")
    for f, (s, n) in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
        print(f"| `{f}` | {100 * s / tot_s:.1f} % | {100 * n / tot_i:.1f} % |")
    print("\n## Top 30 lines by samples\n\n| location | samples | instructions | main opcodes | source |\n|---|---|---|---|---|")
    for loc, (s, n, ops) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:30]:
        top = ", ".join(f"{o} {100 * c / max(n, 1):.0f}%" for o, c in ops.most_common(3))
        print(f"| `{loc[0]}:{loc[1]}` | {100 * s / tot_s:.1f} % | {100 * n / tot_i:.1f} % | {top} | `{src(loc)}` |")


if __name__ == "__main__":
    main()
