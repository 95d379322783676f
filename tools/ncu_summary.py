#!/usr/bin/env python
"""Compact summary of an .ncu-rep (read on the CPU box with `ncu -i`): per captured launch the duration, DRAM bytes,
throughput fractions, occupancy, issue activity, shared-memory wavefronts and the warp-stall sample counts; plus the
hottest SASS lines by stall samples.  python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/NAME"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_op_red.sum",
        "lts__t_sectors_op_atom.sum", "smsp__inst_executed.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
        for i, h in enumerate(hdr):
            if h in KEYS or h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("_not_issued"):
                try:
                    d[h + (" [%s]" % units[i] if units[i] else "")] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        launches.append(d)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    hot = []
    if len(srows) > 2 and "# Samples" in srows[1]:
        h2 = srows[1]
        si = h2.index("# Samples")
        data = [r for r in srows[2:] if len(r) > si and r[si].isdigit()]
        tot = sum(int(r[si]) for r in data) or 1
        for k, r in sorted(enumerate(data), key=lambda x: -int(x[1][si]))[:15]:
            hot.append({"pct_of_samples": round(100.0 * int(r[si]) / tot, 2), "sass": r[1].strip(), "previous": data[k - 1][1].strip() if k else ""})
    json.dump({"report": rep, "launches": launches, "hottest_sass_first_launch": hot}, open(out + ".json", "w"), indent=1)
    print("wrote", out + ".json", len(launches), "launch(es)")


if __name__ == "__main__":
    main()
