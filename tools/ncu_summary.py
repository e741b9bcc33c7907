"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean, share."""
import collections, csv, re, sys

def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    m = re.match(r"(?:void )?([\w:]+)", name)
    base = m.group(1) if m else name
    t = re.search(r"<([^()]*?)>\(", name)
    return base + ("<" + t.group(1) + ">" if t and base.startswith("clica") else "")

def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000.0 if row["Metric Unit"] == "ns" else (v * 1000.0 if row["Metric Unit"] == "ms" else v)
        k = short(row["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us total (cold-cache, serialised: compare SHARES)")
    print("| kernel | launches | total us | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} | {100 * v[1] / tot:.1f}% |")

if __name__ == "__main__":
    main(sys.argv[1])
