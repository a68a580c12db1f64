"""One CSV row per captured kernel launch from an .ncu-rep (raw page): the numbers DESIGN.md / bench.py cite.
python tools/ncu_table.py report.ncu-rep > profiles/xyz.csv"""
import csv
import subprocess
import sys

COLS = [
    ("duration_us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"),
    ("dyn_smem_B", "launch__shared_mem_per_block_dynamic"),
    ("dram_read_B", "dram__bytes_read.sum"),
    ("dram_write_B", "dram__bytes_write.sum"),
    ("dram_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_active_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1tex_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps_active", "sm__warps_active.avg.per_cycle_active"),
    ("sm_cycles", "sm__cycles_elapsed.max"),
    ("inst_executed", "smsp__inst_executed.sum"),
]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ucol = dict(zip(hdr, units))
    w = csv.writer(sys.stdout)
    w.writerow(["id", "kernel"] + [c for c, _ in COLS])
    for i, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "")
        name = name.replace("void unnamed>::", "").replace("sparkcodec::", "")
        name = name.split("(")[0] if "<" not in name else name[:name.index(">") + 1]
        vals = []
        for c, k in COLS:
            v = d.get(k, "")
            try:
                f = float(v)
                f *= UNIT.get(ucol.get(k, ""), 1.0) if c.endswith("_B") or c == "duration_us" else 1.0
                vals.append(f"{f:.6g}")
            except ValueError:
                vals.append(v)
        w.writerow([i, name] + vals)


if __name__ == "__main__":
    main()
