"""Summaries for profiles/: (1) launch list CSV (ncu --metrics gpu__time_duration.sum,dram__bytes_*)
-> per-kernel share of the step; (2) .ncu-rep of a --set full capture -> key metrics per launch."""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    count = collections.Counter()
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            val = val / 1e3 if unit in ("ns", "nsecond") else (val * 1e3 if unit in ("ms", "msecond") else val)  # -> us
            count[name] += 1
        elif unit in ("Kbyte",):
            val *= 1e3
        elif unit in ("Mbyte",):
            val *= 1e6
        elif unit in ("Gbyte",):
            val *= 1e9
        per[name][m] += val
    tot = sum(v["gpu__time_duration.sum"] for v in per.values())
    print(f"| kernel | launches | total us | share | dram read GB | dram write GB |")
    print("|---|---|---|---|---|---|")
    for name, v in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        t = v["gpu__time_duration.sum"]
        print(f"| `{name}` | {count[name]} | {t:.0f} | {100 * t / tot:.1f}% | {v['dram__bytes_read.sum'] / 1e9:.2f} | "
              f"{v['dram__bytes_write.sum'] / 1e9:.2f} |")


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |")
    print("|---|---|" + "---|" * (len(rows) - 2))
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"| `{w}` | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |")


def traffic(path, algorithmic_bytes_per_launch, out_json, note):
    """profiles/traffic.json: measured DRAM bytes per algorithmic byte of the captured launches."""
    import json

    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    tot = []
    for r in rows[2:]:
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            b += float(r[i].replace(",", "")) * scale[units[i]]
        tot.append(b)
    ratio = sum(tot) / (len(tot) * algorithmic_bytes_per_launch)
    doc = {"dram_bytes_per_algorithmic_byte": ratio, "dram_bytes_per_launch": tot,
           "algorithmic_bytes_per_launch": algorithmic_bytes_per_launch,
           "source": f"ncu --set full, {len(tot)} launches of qls_rs_pass_kernel, {note} ({path.split('/')[-1]})"}
    with open(out_json, "w") as fh:
        json.dump(doc, fh, indent=1)
    print(json.dumps(doc))


if __name__ == "__main__":
    if sys.argv[1] == "--traffic":  # --traffic report.ncu-rep <algorithmic bytes per launch> out.json "note"
        traffic(sys.argv[2], float(sys.argv[3]), sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else "")
    elif sys.argv[1].endswith(".csv"):
        launches(sys.argv[1])
    else:
        report(sys.argv[1])
