"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the small tracked summaries under profiles/.

  python scripts/ncu_summary.py <tag> <launches.csv> <report.ncu-rep> "<comment>"
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    tag, launches, rep, comment = sys.argv[1:5]
    if launches != "-":
        rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
        with open(f"profiles/{tag}_launches.csv", "w") as f:
            f.write(f"# {comment}\n# ncu --metrics gpu__time_duration.sum --clock-control none (per-launch times are cold-cache and serialised)\n")
            f.write("id,kernel,block,grid,gpu__time_duration_ns\n")
            for r in rows:
                f.write(f"{r[0]},{r[4].split('(')[0].replace('void ', '')},{r[7]},{r[8]},{r[-1]}\n")
    if rep != "-":
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(io.StringIO(out)))
        hdr, unit = rr[0], rr[1]
        with open(f"profiles/{tag}_full.txt", "w") as f:
            f.write(f"# {comment}\n# ncu --set full --clock-control none --import-source on (one launch)\n")
            for val in rr[2:]:
                f.write(f"## kernel: {val[hdr.index('Kernel Name')]}\n")
                for k in WANT:
                    if k in hdr:
                        i = hdr.index(k)
                        f.write(f"{k} [{unit[i]}] = {val[i]}\n")


if __name__ == "__main__":
    main()
