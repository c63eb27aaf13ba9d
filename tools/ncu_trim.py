"""Trim an `ncu -i X.ncu-rep --page raw --csv` dump to the columns the profiles/ summaries keep.  Usage: python tools/ncu_trim.py raw.csv out.csv "source note" """
import csv
import sys

COLS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def main(raw, out, note):
    with open(raw) as f:
        rows = [r for r in csv.reader(l for l in f if not l.startswith("=="))]
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in COLS if c in ix]
    with open(out, "w") as f:
        f.write(f'"# {note}"\n')
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + cols)
        w.writerow([""] + [units[ix[c]] for c in cols])
        for r in data:
            w.writerow([r[ix["Kernel Name"]][:90]] + [r[ix[c]] for c in cols])
    print(open(out).read()[:3000])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
