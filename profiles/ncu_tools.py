"""Helpers to read .ncu-rep files on the CPU box (ncu -i ... --page raw/source --csv)."""
import csv
import io
import subprocess
import sys


def raw_metrics(rep, wanted=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if wanted is None or any(h.startswith(w) for w in wanted):
                d[h] = (v, u)
        res.append(d)
    return res


def source_hot_lines(rep, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdrs = [i for i, r in enumerate(rows) if "Instructions Executed" in r]
    if not hdrs:
        return [], 0, 0
    h = rows[hdrs[0]]
    ie, smp = h.index("Instructions Executed"), h.index("# Samples")
    agg, cur = {}, None
    for r in rows:
        if r and r[0] == "File Name":
            cur = r[1].split("/")[-1]
        elif len(r) > max(ie, smp) and r[0].isdigit() and r[ie].replace(".", "").isdigit():
            key = (cur, int(r[0]))
            a = agg.setdefault(key, [r[1], 0.0, 0.0])
            a[1] += float(r[ie])
            a[2] += float(r[smp]) if r[smp].replace(".", "").isdigit() else 0.0
    items = [(k[0], k[1], v[0], v[1], v[2]) for k, v in agg.items()]
    tot, ts = sum(i[3] for i in items), sum(i[4] for i in items)
    items.sort(key=lambda o: -o[4])
    return items[:top], tot, ts


WANTED = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
          "launch__occupancy_limit", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "smsp__average_warps_issue_stalled_barrier_per_issue_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active",
          "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active",
          "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active", "smsp__average_warps_issue_stalled_wait_per_issue_active",
          "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active", "sm__inst_executed.sum", "lts__t_bytes.sum",
          "l1tex__t_bytes.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg"]

if __name__ == "__main__":
    rep = sys.argv[1]
    for d in raw_metrics(rep, WANTED):
        for k, (v, u) in d.items():
            if ".pct_of_peak" in k and not k.startswith(("gpu__dram", "sm__throughput", "sm__warps", "smsp__issue")):
                continue
            print(f"{k:90s} {v:>20s} {u}")
    items, tot, ts = source_hot_lines(rep, int(sys.argv[2]) if len(sys.argv) > 2 else 40)
    print(f"--- hottest source lines (total inst {tot:.3g}, samples {ts:.3g})")
    for f, ln, src, inst, s in items:
        print(f"{f}:{ln:4d} inst%={100*inst/max(tot,1):5.2f} smp%={100*s/max(ts,1):5.2f} | {src[:110]}")
