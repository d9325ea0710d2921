"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
per = []
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    unit = row["Metric Unit"]
    us = v / 1000 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000)
    short = re.sub(r"\(.*", "", row["Kernel Name"])
    short = re.sub(r"void |\(anonymous namespace\)::|<unnamed>::", "", short)
    agg[short][0] += 1
    agg[short][1] += us
    tot += us
    per.append((us, short, row.get("Grid Size", ""), row.get("Block Size", "")))
print(f"total {tot:.1f} us over {len(per)} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  n={n:4d} avg={t / n:8.1f}  {k[:100]}")
if "--single" in sys.argv:
    for us, short, g, b in sorted(per, reverse=True)[:25]:
        print(f"{us:9.1f} {short[:80]} {g} {b}")
