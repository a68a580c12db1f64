"""Summarise an .ncu-rep: one block of key metrics per captured kernel (raw page -> compact text).
python tools/ncu_summary.py report.ncu-rep [extra-metric-substring ...]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex.sum", "smsp__inst_executed.sum",
    "sm__warps_active.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled",
    "launch__registers_per_thread", "launch__grid_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_xu", "sm__pipe_fma_cycles_active",
    "sm__pipe_alu_cycles_active", "sm__inst_executed_pipe_lsu",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", d.get("Kernel Name", "")[:110])
        for k, u in zip(hdr, units):
            if any(k.startswith(p) or p in k for p in KEYS + extra) and "TriageCompute" not in k and "Triage" not in k:
                if ".max" in k and "cycles_elapsed" not in k or ".min" in k or ".sum.pct" in k or ".sum.per_second" in k:
                    continue
                print(f"   {k:90s} {d[k]:>16s} {u}")


if __name__ == "__main__":
    main()
