"""Summarise .ncu-rep files (ncu -i ... --page raw --csv) into a small text table for profiles/."""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active"]
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]; units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [k for k in hdr if any(k == key or k.startswith(key) for key in KEYS)]
    print(f"# {path}")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]][:60]
        print(f"- {name}  grid={r[idx.get('Grid Size', 0)] if 'Grid Size' in idx else ''} block={r[idx['Block Size']] if 'Block Size' in idx else ''}")
        for c in cols:
            print(f"    {c:75s} {r[idx[c]]:>18s} {units[idx[c]]}")
if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
