"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals."""
import collections, csv, re, sys
path, nfwd = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("nsecond", "ns") else v * 1e3 if u in ("msecond", "ms") else v
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"total {tot / nfwd:.1f} us per forward ({nfwd:g} forwards)")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:22]:
    print(f"{t / nfwd:9.1f} us/fwd {100 * t / tot:5.1f}%  n/fwd={n / nfwd:6.1f}  mean {t / n:8.2f} us  {k[:80]}")
