"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few numbers the roofline needs."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active"]

path = sys.argv[1]
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [k for k in KEYS if k in idx]
tens = [h for h in hdr if "tensor" in h and h not in cols][:6]
print("# " + path)
print("columns:", ", ".join(f"{c} [{units[idx[c]]}]" for c in cols + tens))
for r in data:
    name = r[idx["Kernel Name"]][:48]
    vals = [r[idx[c]] for c in cols + tens]
    print(f"{r[idx['ID']]:>3} {name:48s} " + " | ".join(vals))
