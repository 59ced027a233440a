#!/usr/bin/env python
"""tools/ncu_summary.py WORKLOAD FILE.ncu-rep [FILE2.ncu-rep ...]

Reads `ncu --set full` reports (with the CPU-side `ncu -i ... --page raw --csv`, no GPU needed) and merges the
counters the roofline discussion uses into profiles/r02_ncu_summary.json:

    { workload: { kernel: { "dram_bytes_per_launch": ..., "duration_us": ..., "inst_executed": ...,
                            "fp32_ops": {"fadd": .., "fmul": .., "ffma": ..}, "pipe_fma_pct_active": .., ... } } }

bench.py takes `roofline.traffic` (and the `ncu` sub-record) for whichever workload it runs from that file.
Several launches of one kernel in a report are averaged."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_ncu_summary.json")
WANT = {
    "duration_us": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "inst_executed": "smsp__inst_executed.sum",
    # (`--set full` reports these three as a rate: thread-level instructions per elapsed cycle, summed over the GPU)
    "fadd_thread_inst_per_cycle": "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
    "fmul_thread_inst_per_cycle": "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
    "ffma_thread_inst_per_cycle": "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "pipe_fma_pct_of_active": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "pipe_fma_pct_of_elapsed": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "pipe_alu_pct_of_active": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "pipe_lsu_pct_of_active": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "pipe_xu_pct_of_active": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp_cycles_active_avg": "smsp__cycles_active.avg",
    "sm_cycles_elapsed_max": "sm__cycles_elapsed.max",
    "warps_active_pct_of_peak": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "shared_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "shared_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "registers_per_thread": "launch__registers_per_thread",
    "grid_size": "launch__grid_size",
    "block_size": "launch__block_size",
    "waves_per_sm": "launch__waves_per_multiprocessor",
    "nvlink_rx_bytes": "nvlrx__bytes.sum",
    "nvlink_tx_bytes": "nvltx__bytes.sum",
}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def short(name):
    m = re.search(r"(\w+)(<|\()", name.replace("gi2d::<unnamed>::", "").replace("void ", ""))
    return m.group(1) if m else name[:40]


def main():
    wl, files = sys.argv[1], sys.argv[2:]
    summ = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for f in files:
        raw = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        acc = {}
        for r in rows[2:]:
            if len(r) < len(hdr):
                continue
            k = short(r[col["Kernel Name"]])
            d = acc.setdefault(k, {"launches": 0})
            d["launches"] += 1
            for key, metric in WANT.items():
                if metric in col and r[col[metric]] not in ("", "n/a"):
                    try:
                        v = float(r[col[metric]].replace(",", "")) * UNIT_SCALE.get(units[col[metric]], 1)
                    except ValueError:
                        continue
                    d[key] = d.get(key, 0.0) + v
        for k, d in acc.items():
            n = d.pop("launches")
            rec = {key: v / n for key, v in d.items()}
            rec["launches_averaged"] = n
            if "dram_read_bytes" in rec:
                rec["dram_bytes_per_launch"] = rec["dram_read_bytes"] + rec.get("dram_write_bytes", 0.0)
            if "smsp_cycles_active_avg" in rec and rec.get("sm_cycles_elapsed_max"):
                rec["cycles_active_over_elapsed"] = rec["smsp_cycles_active_avg"] / rec["sm_cycles_elapsed_max"]
            rec["source"] = "profiles/" + os.path.basename(f).replace(".ncu-rep", "_raw.csv")
            summ.setdefault(wl, {})[k] = rec
        # keep the raw page next to the summary (the .ncu-rep itself is too big for the repo)
        open(os.path.join(ROOT, "profiles", os.path.basename(f).replace(".ncu-rep", "_raw.csv")), "w").write(raw)
    json.dump(summ, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, {w: sorted(v) for w, v in summ.items()})


if __name__ == "__main__":
    main()
