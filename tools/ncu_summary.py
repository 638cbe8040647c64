"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few numbers the roofline needs.
Usage: python tools/ncu_summary.py report.ncu-rep > profiles/rNN_ncu_full_<tag>.txt"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("lts__t_bytes.sum", "l2_MB"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%act"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_%el"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_act_%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dsmem")]

path = sys.argv[1]
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [(k, n) for k, n in COLS if k in idx]
print(f"# {path}  (ncu --set full --clock-control none; per-launch, cold-ish cache, serialised)")
print("# units: " + ", ".join(f"{n}[{units[idx[k]]}]" for k, n in cols))
print(f"{'id':>3} {'kernel':44s} " + " ".join(f"{n:>16s}" for _, n in cols))
for r in data:
    name = r[idx["Kernel Name"]].split("(")[0][:44]
    vals = []
    for k, _ in cols:
        v = r[idx[k]].replace(",", "")
        try:
            vals.append(f"{float(v):16.3f}")
        except ValueError:
            vals.append(f"{v:>16s}")
    print(f"{r[idx['ID']]:>3} {name:44s} " + " ".join(vals))
