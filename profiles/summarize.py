"""Turn the ncu artefacts of a gpurun call into the committed per-round summary.

    python profiles/summarize.py r01 gpurun_out/launches_r01.csv gpurun_out/hmc_step_r01.ncu-rep \
        [gpurun_out/bench.json]

writes profiles/<round>_launches.csv (the `--metrics gpu__time_duration.sum` launch list of
profiles/profile_chain.py, one warm-up + one timed chain) and profiles/<round>_summary.md.
"""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def launch_list(path):
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            val = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1}.get(unit, 1)
            rows.append((int(r["ID"]), r["Kernel Name"].split("(")[0][:80], ns))
    return rows


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def stall_breakdown(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > ci["# Samples"]]
    tot = sum(float(r[ci["# Samples"]] or 0) for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    t = {h[6:]: sum(float(r[ci[h]] or 0) for r in data) / tot for h in stalls}
    ops = collections.Counter()
    inst = 0.0
    for r in data:
        txt = r[ci["Source"]].strip()
        o = (txt.split()[1] if txt.startswith("@") else txt.split()[0]).split(".")[0]
        n = float(r[ci["Instructions Executed"]] or 0)
        ops[o] += n
        inst += n
    return t, {k: v / inst for k, v in ops.most_common(8)}


def main():
    rnd, launches_csv, rep = sys.argv[1:4]
    bench = json.load(open(sys.argv[4])) if len(sys.argv) > 4 else None
    rows = launch_list(launches_csv)
    with open(os.path.join(HERE, f"{rnd}_launches.csv"), "w") as fh:
        fh.write("id,kernel,duration_ns\n")
        for i, k, ns in rows:
            fh.write(f"{i},{k},{ns:.0f}\n")
    # second half = the timed chain (first half is the warm-up chain incl. weight packing)
    mine = [r for r in rows if r[1].startswith(("k_", "void k_"))]
    half = mine[len(mine) // 2:]
    by = collections.defaultdict(lambda: [0, 0.0])
    for _, k, ns in half:
        name = k.replace("void ", "")
        by[name][0] += 1
        by[name][1] += ns
    tot = sum(v[1] for v in by.values())
    m, units = raw_metrics(rep)
    stalls, ops = stall_breakdown(rep)
    g = lambda k: m.get(k, "n/a")
    lines = [f"# {rnd}: ncu summary (B200, profiles/profile_chain.py = one BASELINE-config-2 AIS call)", ""]
    lines += ["## Launch list of one chain (cold-cache, serialised under ncu: compare shares)", "",
              "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, (c, ns) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f} % |")
    lines += ["", f"own kernels per chain: {sum(v[0] for v in by.values())}; total {tot / 1e6:.2f} ms under ncu", ""]
    lines += ["## dominant kernel, full capture (`--set full --clock-control none`)", ""]
    dr, dw = float(g("dram__bytes_read.sum")), float(g("dram__bytes_write.sum"))
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}
    dr *= scale.get(units.get("dram__bytes_read.sum"), 1)
    dw *= scale.get(units.get("dram__bytes_write.sum"), 1)
    lines += [f"* duration {g('gpu__time_duration.sum')} {units.get('gpu__time_duration.sum')}, grid {g('launch__grid_size')}, "
              f"{g('launch__registers_per_thread')} regs/thread, dynamic smem {g('launch__shared_mem_per_block_dynamic')} KB",
              f"* DRAM traffic per launch: read {dr / 1e6:.2f} MB + write {dw / 1e6:.2f} MB = **{(dr + dw) / 1e6:.2f} MB** "
              f"(algorithmic Point+noise bytes: 2048 x 924 B = 1.89 MB; the 10.2 MB weight blob is read once then L2-resident)",
              f"* L2: tex read sectors {g('lts__t_sectors_srcunit_tex_op_read.sum')} (x32 B), hit rate {g('lts__t_sector_hit_rate.pct')} %",
              f"* FMA pipe active {g('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active')} %, issue slots busy "
              f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active')} %, warps active {g('sm__warps_active.avg.pct_of_peak_sustained_active')} % of max",
              f"* tensor pipe ({(bench or {}).get('roofline', {}).get('pipe', 'see bench line')}): active {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} % of the active SMs' cycles; "
              f"ALU pipe {g('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active')} %; SM clock during the capture "
              f"{g('sm__cycles_elapsed.avg.per_second')} {units.get('sm__cycles_elapsed.avg.per_second')}",
              f"* instructions {g('smsp__inst_executed.sum')}; mix: " + ", ".join(f"{k} {100 * v:.1f} %" for k, v in ops.items()),
              "* warp-stall samples: " + ", ".join(f"{k} {100 * v:.1f} %" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:9]), ""]
    if bench:
        r = bench["roofline"]
        lines += ["## bench.py line of the same build", "",
                  f"* value {bench['value']:.0f} particles/s ({bench['ms_per_step']:.2f} ms/step), e2e {bench['e2e']['value']:.0f} particles/s",
                  f"* `k_hmc_step` {r['kernel_ms']:.3f} ms/launch (CUDA events) -> {r['achieved']:.2f} TFLOP/s algorithmic = "
                  f"{100 * r['pipe_frac']:.1f} % of the engine's own ceiling on the SMs this batch occupies ({r['pipe_peak']:.1f} TF, see `pipe` in the bench line), "
                  f"{100 * r['frac']:.2f} % of the measured bf16 cuBLAS peak",
                  f"* kernel share of the step: {100 * bench['config']['M'] * r['kernel_ms'] / bench['ms_per_step']:.1f} % (events) vs "
                  f"{100 * sum(v[1] for k, v in by.items() if k.startswith('k_hmc_step')) / tot if tot else 0:.1f} % (ncu launch list)",
                  f"* clocks: {bench.get('clocks')}", ""]
        if "cpu_baseline" in bench:
            lines += [f"* cpu_baseline: {bench['cpu_baseline']}", ""]
        if "parity" in bench:
            lines += [f"* parity (CUDA vs CPU port, identical injected noise): {bench['parity']}", ""]
    json.dump({"round": rnd, "kernel": "k_hmc_step", "dram_bytes_per_launch": dr + dw,
               "duration_ms_under_ncu": float(g("gpu__time_duration.sum")),
               "fma_pipe_active_pct": float(g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")),
               "tensor_pipe_active_pct": float(g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")),
               "issue_active_pct": float(g("smsp__issue_active.avg.pct_of_peak_sustained_active"))},
              open(os.path.join(HERE, f"{rnd}_metrics.json"), "w"), indent=1)
    open(os.path.join(HERE, f"{rnd}_summary.md"), "w").write("\n".join(lines))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
