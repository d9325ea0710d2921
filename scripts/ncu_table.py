"""Markdown table + traffic JSON from .ncu-rep files, one row per kernel name: launches captured, mean and largest
duration, DRAM bytes read / written per launch (mean; the JSON holds read + write of the LARGEST launch of each kernel,
i.e. its stage-0 instance — the one bench.py times), DRAM %, tensor-pipe %, issue-slot %, registers. A kernel that appears
in several reports keeps the row of the first report that has it.

    python scripts/ncu_table.py out.md traffic.json rep1.ncu-rep [rep2 ...]
"""
import collections
import csv
import json
import re
import subprocess
import sys

M = {"t": "gpu__time_duration.sum", "r": "dram__bytes_read.sum", "w": "dram__bytes_write.sum",
     "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
     "grid": "launch__grid_size", "block": "launch__block_size"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6, "msecond": 1e3, "usecond": 1.0}


def main():
    out_md, out_json, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    agg = collections.OrderedDict()
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            continue
        h, u = rows[0], rows[1]
        col = {k: h.index(v) for k, v in M.items() if v in h}
        ik = h.index("Kernel Name")
        for r in rows[2:]:
            name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|^void ", "", r[ik])
            name = re.sub(r"\(.*$", "", name)
            if name in agg and agg[name]["rep"] != [rep.split("/")[-1]]:
                continue
            d = agg.setdefault(name, collections.defaultdict(list))
            for k, i in col.items():
                try:
                    d[k].append(float(r[i].replace(",", "")) * SCALE.get(u[i], 1.0))
                except ValueError:
                    pass
            d["rep"] = [rep.split("/")[-1]]
    mean = lambda v: sum(v) / len(v) if v else float("nan")
    lines = ["| kernel | launches | us (mean) | us (max) | DRAM read MB | DRAM write MB | DRAM % | tensor pipe % | issue % | regs | grid x block | capture |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    traffic = {}
    for name, d in agg.items():
        lines.append(f"| `{name}` | {len(d['t'])} | {mean(d['t']):.1f} | {max(d['t']):.1f} | {mean(d['r']) / 1e6:.2f} | {mean(d['w']) / 1e6:.2f} | "
                     f"{mean(d['dram']):.1f} | {mean(d['tensor']):.1f} | {mean(d['issue']):.1f} | {mean(d['regs']):.0f} | "
                     f"{mean(d['grid']):.0f} x {mean(d['block']):.0f} | {d['rep'][0]} |")
        big = max(a + b for a, b in zip(d["r"], d["w"])) if d["r"] and d["w"] else None
        key = name.split("<")[0]
        traffic[key] = max(traffic.get(key) or 0, big or 0) or None
        traffic[name] = big
    open(out_md, "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(out_json, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
