"""Summarise an .ncu-rep: key metrics + hottest SASS lines (by samples / instructions)."""
import csv
import io
import subprocess
import sys


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main(rep, top=28):
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
            "launch__occupancy_limit_registers", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__lsu_writeback_active_mem_lg.sum.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
    for h, u, v in zip(hdr, units, vals):
        if h in keys or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.3):
            print(f"{h},{u},{v}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    print("kernel:", src[0][1][:150])
    hdr = src[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = src[2:]
    ti = sum(int(r[ix["Instructions Executed"]]) for r in data)
    ts = sum(int(r[ix["# Samples"]]) for r in data)
    print("total warp-inst", ti, "samples", ts)
    order = sorted(range(len(data)), key=lambda n: -int(data[n][ix["# Samples"]]))[:top]
    for n in sorted(order):
        r = data[n]
        print(f"{n:5d} {r[ix['Source']][:58]:58s} inst={100 * int(r[ix['Instructions Executed']]) / ti:5.2f}% samp={100 * int(r[ix['# Samples']]) / ts:5.2f}% "
              f"long={r[ix['stall_long_sb']]} short={r[ix['stall_short_sb']]} mio={r[ix['stall_mio']]} wait={r[ix['stall_wait']]} bar={r[ix['stall_barrier']]} lg={r[ix['stall_lg']]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 28)
