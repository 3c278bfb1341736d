#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  launch list  (ncu --metrics gpu__time_duration.sum --csv)  -> per-kernel totals and shares
  full capture (ncu --set full ... -o X.ncu-rep)              -> selected raw metrics per launch
"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row.get("Metric Unit", "ns")
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        name = row["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    with open(out, "w") as f:
        f.write("kernel,launches,total_ms,share_pct\n")
        for k, v in tot.most_common():
            f.write("%s,%d,%.3f,%.2f\n" % (k, cnt[k], v / 1e6, 100 * v / T))
        f.write("TOTAL,%d,%.3f,100.00\n" % (sum(cnt.values()), T / 1e6))
    print(open(out).read())


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("metric,unit," + ",".join("launch%d" % i for i in range(len(rows) - 2)) + "\n")
        for i, h in enumerate(hdr):
            if h in WANT or h in ("Kernel Name",):
                f.write("%s,%s,%s\n" % (h, units[i], ",".join('"%s"' % r[i].split("(")[0] for r in rows[2:])))
    print(open(out).read())


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    (launches if mode == "launches" else full)(src, dst)
