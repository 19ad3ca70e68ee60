#!/usr/bin/env python
"""profiles/kernel_inst_counts.json from an `ncu --set full` capture of scripts/prof_dense.py:

    ncu --set full --metrics smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_xu.sum \\
        --clock-control none --import-source on -k regex:'p2p_moment|dense_f2|dense_pass|wide_tc|wide_pass' -o gpurun_out/r2_configs \\
        python scripts/prof_dense.py
    python scripts/ncu_inst_counts.py gpurun_out/r2_configs.ncu-rep

Per kernel: warp instructions per residual (smsp__inst_executed.sum / n), FMA-pipe warp instructions per residual,
DRAM bytes per launch, duration and the headline percentages; the LAST launch of each kernel in the capture is used.
bench.py multiplies the per-residual counts by residuals per second for its issue-slot rooflines."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = [  # key, substring of the kernel name, residuals per launch
    ("p2p_gen2_f32", "p2p_moment2_kernel", 100_000_000),
    ("camera6_central_f32", "dense_f2_kernel<PinholeModel", 50_000_000),
    ("camera15_central_f32", "wide_tc_kernel<PinholeDistortModel", 50_000_000),
    ("curve_central_f32", "dense_f2_kernel<ExpCurveModel", 10_000_000),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "inst": 1.0, "": 1.0,
         "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}  # times in microseconds


def main():
    rep = sys.argv[1]  # an .ncu-rep, or the output of `ncu -i <rep> --page raw --csv` (for captures too large to bring back)
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def val(r, name, default=None):
        if name not in hdr:
            return default
        i = hdr.index(name)
        try:
            return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
        except ValueError:
            return default

    out = {}
    for key, sub, n in KERNELS:
        hits = [r for r in data if sub in r[hdr.index("Kernel Name")]]
        if not hits:
            continue
        r = hits[-1]
        inst = val(r, "smsp__inst_executed.sum")
        fma = val(r, "smsp__inst_executed_pipe_fma.sum")
        out[key] = {
            "kernel": r[hdr.index("Kernel Name")][:120], "n": n,
            "warp_inst_per_residual": inst / n if inst else None,
            "fma_pipe_inst_per_residual": fma / n if fma else None,
            "thread_inst_per_residual": inst * 32 / n if inst else None,
            "dram_bytes_per_launch": (val(r, "dram__bytes_read.sum", 0.0) or 0.0) + (val(r, "dram__bytes_write.sum", 0.0) or 0.0),
            "duration_us": val(r, "gpu__time_duration.sum"),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "tensor_pipe_inst_per_residual": (val(r, "smsp__inst_executed_pipe_tensor.sum") or 0.0) / n,
            "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "registers": val(r, "launch__registers_per_thread"),
            "shared_bank_conflicts": val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
            "shared_wavefronts": val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
            "source": "profiles/" + os.path.basename(rep).replace(".ncu-rep", "").replace("_rawpage.csv", "") + "_ncu_raw.csv",
        }
    path = os.path.join(ROOT, "profiles", "kernel_inst_counts.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    # keep a trimmed copy of the raw page next to it (the .ncu-rep itself is too large to commit)
    keep = [i for i, h in enumerate(hdr) if any(k in h for k in (
        "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration", "smsp__issue_active", "smsp__inst_executed",
        "dram__bytes", "dram__throughput", "gpu__dram_throughput", "dram__cycles_active", "sm__pipe_", "launch__registers",
        "launch__occupancy", "sm__warps_active", "issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared", "sm__cycles_elapsed.avg.per_second", "smsp__warps_eligible",
        "sm__throughput", "sm__inst_executed_pipe", "lts__t_sector_hit_rate"))]
    with open(os.path.join(ROOT, out[next(iter(out))]["source"]) if out else os.devnull, "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [r[hdr.index("Kernel Name")][:100] for r in data])
        for i in keep:
            w.writerow([hdr[i], units[i]] + [r[i] for r in data])
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
