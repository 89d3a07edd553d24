"""Turns ncu output into the small JSON summaries kept under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.json> "<command>"   # --metrics gpu__time_duration.sum --csv log
  python tools/ncu_summary.py full <report.ncu-rep> <out.json> "<what>"        # one --set full capture (reads it with ncu -i)
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed.sum.per_cycle_active", "sm__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio")


def short(name):
    name = re.sub(r"\(bool\)|\(int\)|mfc::|void ", "", name)
    return re.sub(r"\(.*\)$", "", name)


def launches(path, out, command):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    iu = h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        us = v / 1000.0 if r[iu] in ("ns", "nsecond") else v
        k = short(r[ik])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    ks = sorted(({"kernel": k, "launches": a[0], "us": round(a[1], 1), "share": round(a[1] / tot, 4)} for k, a in agg.items()),
                key=lambda d: -d["us"])
    json.dump({"command": command, "note": "cold-cache, serialised launches: compare shares, not absolutes", "total_us": round(tot, 1),
               "kernels": ks}, open(out, "w"), indent=1)
    print(json.dumps(ks[:6]))


def full(rep, out, what):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units, vals = rows[0], rows[1], rows[2]
    d = {"Kernel Name": short(vals[h.index("Kernel Name")]), "Block Size": vals[h.index("Block Size")], "Grid Size": vals[h.index("Grid Size")]}
    for k in KEEP:
        if k in h:
            i = h.index(k)
            d[k] = "%s %s" % (vals[i], units[i])
    d["_what"] = what
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d)[:600])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
