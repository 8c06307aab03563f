#!/usr/bin/env python
"""Turn ncu artefacts from gpurun_out/ into the small text summaries committed under profiles/.

    python tools/ncu_summary.py report  gpurun_out/prof.ncu-rep   profiles/r1_full_n32768.md
    python tools/ncu_summary.py launches gpurun_out/launches.csv  profiles/r1_launches.md

(`ncu -i` works without a GPU.)"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.per_cycle_active", "warps active / SM"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe cycles active %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe instruction issue %"),
    ("derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2", "DFMA thread instructions x2 (flop)"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread instructions"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread instructions"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe (MUFU.RCP64H) %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM read bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts %"),
]
STALLS = ["barrier", "wait", "short_scoreboard", "long_scoreboard", "math_pipe_throttle", "not_selected", "selected",
          "branch_resolving", "no_instruction", "dispatch_stall", "mio_throttle", "lg_throttle", "membar", "sleeping"]


def report(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none: {rep.split('/')[-1]}\n\n")
        f.write("Per-launch values (one column per profiled launch).  Times under ncu are cold-cache and serialised;\n"
                "bench.py's CUDA-event numbers are the ones to quote.\n\n")
        names = [r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("ah::", "") for r in data]
        f.write("| metric | unit | " + " | ".join(names) + " |\n|---|---|" + "---|" * len(names) + "\n")
        for key, label in KEYS:
            if key in ix:
                f.write(f"| {label} (`{key}`) | {units[ix[key]]} | " + " | ".join(r[ix[key]] for r in data) + " |\n")
        f.write("\nWarp stall reasons (warps stalled per issued instruction, `smsp__average_warps_issue_stalled_*_per_issue_active.ratio`):\n\n")
        f.write("| stall | " + " | ".join(names) + " |\n|---|" + "---|" * len(names) + "\n")
        for s in STALLS:
            key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if key in ix:
                f.write(f"| {s} | " + " | ".join(f"{float(r[ix[key]]):.3f}" for r in data) + " |\n")


def launches(csvfile, out):
    lines = [l for l in open(csvfile) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        key = (name, r["Grid Size"], r["Block Size"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v_ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v_ms
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({csvfile.split('/')[-1]}): `--metrics gpu__time_duration.sum --clock-control none`\n\n")
        f.write(f"{len(rows)} launches, {tot:.2f} ms of kernel time in total (serialised, cold cache: compare shares).\n\n")
        f.write("| kernel | grid | block | launches | total ms | mean ms | share |\n|---|---|---|---|---|---|---|\n")
        for (name, g, b), (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {g} | {b} | {c} | {ms:.3f} | {ms / c:.3f} | {100 * ms / tot:.1f} % |\n")


if __name__ == "__main__":
    {"report": report, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
