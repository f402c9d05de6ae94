"""Summarise an .ncu-rep (brought back from the GPU box in gpurun_out/) into a small JSON under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/r01_X_ncu_summary.json "free-text note" """
import csv
import io
import json
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__cluster_size",
        "launch__grid_size", "launch__block_size", "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "Kernel Name", "Grid Size", "Block Size")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        kernels.append({h: [v, u] for h, u, v in zip(hdr, units, vals) if h in KEEP})
    json.dump({"source": rep, "note": note, "kernels": kernels}, open(out, "w"), indent=1)
    for k in kernels:
        print(json.dumps(k)[:600])


if __name__ == "__main__":
    main()
