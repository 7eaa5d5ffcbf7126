"""Summarise an .ncu-rep (ncu --set full) into a markdown table for profiles/."""
import csv, subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_thr"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb")]
idx = [(n, hdr.index(k)) for k, n in want if k in hdr]
units = rows[1]
with open(out, "w") as f:
    f.write(f"# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none); units: "
            + ", ".join(f"{n}={units[i]}" for n, i in idx if units[i]) + "\n\n")
    f.write("| " + " | ".join(n for n, _ in idx) + " |\n|" + "---|" * len(idx) + "\n")
    for r in rows[2:]:
        cells = []
        for n, i in idx:
            v = r[i]
            if n == "kernel":
                v = v.split("(")[0].replace("<unnamed>::", "")[:40]
            else:
                try:
                    v = f"{float(v.replace(',', '')):.3f}".rstrip("0").rstrip(".")
                except ValueError:
                    pass
            cells.append(v)
        f.write("| " + " | ".join(cells) + " |\n")
print(open(out).read())
