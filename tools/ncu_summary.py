#!/usr/bin/env python
"""One line per profiled kernel from an .ncu-rep (`ncu --set full`): duration, DRAM bytes, issue / warp activity,
registers, instructions.  Also prints {"kernel": dram bytes per launch} for profiles/traffic.json.
usage: ncu_summary.py report.ncu-rep out.csv [traffic.json]"""
import csv, io, json, re, subprocess, sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    idx = [h.index(c) for c in COLS]
    last = {}
    for r in rows[2:]:
        name = re.sub(r"^void\s+", "", r[ki])
        name = re.sub(r"unnamed>::", "", name).split("(")[0]
        last[name] = r  # the last launch of a kernel is the warm one
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + COLS)
        w.writerow([""] + [units[i] for i in idx])
        for name, r in last.items():
            w.writerow([name] + [r[i] for i in idx])
    if len(sys.argv) > 3:
        ri, wi = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tr = {}
        for name, r in last.items():
            key = name
            if name.startswith("k_dec_rows<0"):
                key = "k_dec_row_lens"
            elif name.startswith("k_dec_rows<1"):
                key = "k_dec_write_rows"
            tr[key] = int(float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]])
        json.dump(tr, open(sys.argv[3], "w"), indent=1)
        print(json.dumps(tr))


if __name__ == "__main__":
    main()
