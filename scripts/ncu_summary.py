"""Summarise ncu reports for profiles/: one CSV row per captured launch with the metrics the roofline
discussion uses.  Usage: python scripts/ncu_summary.py out.csv "header comment" rep1.ncu-rep [rep2 ...]"""
import csv
import io
import subprocess
import sys

COLS = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]


def rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    head, units, body = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(head)}
    have = [c for c in COLS if c in idx]
    yield have, [units[idx[c]] for c in have]
    for r in body:
        yield None, [r[idx[c]] for c in have]


def main():
    out, comment, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        f.write('"# ' + comment.replace('"', "'") + '"\n')
        first = True
        for rep in reps:
            for have, vals in rows(rep):
                if have is not None:
                    if first:
                        w.writerow(have)
                        w.writerow(vals)
                        first = False
                    continue
                w.writerow(vals)


if __name__ == "__main__":
    main()
