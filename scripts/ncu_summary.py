#!/usr/bin/env python
"""Summarises an `ncu --set full` report (run here, no GPU needed: `ncu -i ... --page raw --csv`) as the markdown
table committed under profiles/ and prints the DRAM bytes per launch that bench.py reports as roofline.traffic.

    python scripts/ncu_summary.py gpurun_out/prof_final.ncu-rep > profiles/r2_ncu_full.md
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    traffic = {}
    print(f"# ncu --set full --clock-control none --import-source on: {path}\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?")
        print(f"## {name}\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                print(f"| {k} | {d[k]} | {u.get(k, '')} |")
        try:
            rd = float(d["dram__bytes_read.sum"]) * SCALE.get(u["dram__bytes_read.sum"], 1.0)
            wr = float(d["dram__bytes_write.sum"]) * SCALE.get(u["dram__bytes_write.sum"], 1.0)
            print(f"| dram traffic per launch | {(rd + wr) / 1e9:.3f} | GB |")
            short = name.split("(")[0].split("<")[0].split("::")[-1].strip()
            traffic[short] = rd + wr
        except Exception:
            pass
        print()
    print("<!-- roofline_traffic: " + json.dumps(traffic) + " -->")


if __name__ == "__main__":
    main(sys.argv[1])
