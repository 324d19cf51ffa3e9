"""Turns ncu outputs into the small tracked files under profiles/.

    python tools/ncu_summarize.py launches gpurun_out/launches_r01f.csv profiles/r01f
    python tools/ncu_summarize.py full gpurun_out/prof_r01f.ncu-rep profiles/r01f

`launches`: per-launch list (--metrics gpu__time_duration.sum --csv) -> <prefix>_launches.csv (copy)
and <prefix>_launch_summary.csv (per kernel: launches, total, average, share).
`full`: .ncu-rep of a --set full capture -> <prefix>_ncu_full_key_metrics.csv (one row per captured
launch, selected metrics) and <prefix>_traffic.json (DRAM bytes of the streaming kernel).
"""
import collections
import csv
import io
import json
import shutil
import subprocess
import sys

KEY = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__inst_executed_op_shared_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def launches(src, prefix):
    rows = list(csv.reader(open(src)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        us = v / 1000.0 if r[ui] in ("ns", "nsecond") else v * (1000.0 if r[ui] in ("ms", "msecond") else 1.0)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += us
    shutil.copyfile(src, prefix + "_launches.csv")
    total = sum(a[1] for a in agg.values())
    with open(prefix + "_launch_summary.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_of_listed_kernels"])
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, f"{us:.1f}", f"{us / n:.1f}", f"{100 * us / total:.1f}%"])


def full(rep, prefix):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, units, data = rows[0], rows[1], rows[2:]
    ki = H.index("Kernel Name")
    cols = [(m, H.index(m)) for m in KEY if m in H]
    with open(prefix + "_ncu_full_key_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + [m for m, _ in cols])
        w.writerow([""] + [units[i] for _, i in cols])
        for r in data:
            w.writerow([r[ki]] + [r[i] for _, i in cols])
    for r in data:
        if "k_stream_tma" in r[ki]:
            def to_bytes(m):
                i = H.index(m)
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
                return float(r[i].replace(",", "")) * scale
            rec = {"workload": "c3", "kernel": r[ki], "source": f"ncu --set full, {prefix}_ncu_full_key_metrics.csv",
                   "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
                   "note": "algorithmic bytes 14 B x 3.6e8 = 5.04e9"}
            json.dump(rec, open(prefix + "_traffic.json", "w"), indent=1)
            break


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
