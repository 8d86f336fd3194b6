"""Summarises `ncu --page raw --csv` and `--page source --csv` exports (key metrics, stall mix, opcode mix)."""
import collections
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(units, vals)))
    for k in KEYS:
        if k in d:
            print(f"{k:75s} {d[k][1]:>18s} {d[k][0]}")
    print("stalls per issue:")
    st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v[1]))
          for h, v in d.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for n, v in sorted(st, key=lambda x: -x[1])[:10]:
        print(f"   {n:28s} {v:.3f}")


def src(path, warps):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    c, s = collections.Counter(), collections.Counter()
    tot = 0
    for r in rows[2:]:
        if len(r) <= ie:
            continue
        op = r[ia].split()
        if not op:
            continue
        o = op[0] if not op[0].startswith("@") else op[1]
        o = ".".join(o.split(".")[:2]) if o.startswith("IMAD") else o.split(".")[0]
        n = int(r[ie])
        c[o] += n
        tot += n
        s[o] += int(r[isamp])
    print(f"static SASS instructions {len(rows) - 2}, warp-instructions per warp {tot / warps / 1e6:.3f} M")
    for k, v in c.most_common(16):
        print(f"   {k:12s} {v / warps / 1e6:8.3f} M/warp {100 * v / tot:5.1f}%   stall samples {s[k]}")


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        src(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 2048)
