"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel time share of ONE train step
(the launches between the first and the second fused-SGD launch).  Usage: python tools/launch_summary.py <csv> [out.md]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    sgd = [i for i, r in enumerate(rows) if "sgd_kernel" in r["Kernel Name"]]
    lo, hi = (sgd[0] + 1, sgd[1] + 1) if len(sgd) >= 2 else (0, len(rows))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[lo:hi]:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("mmpl::<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += float(r["Metric Value"].replace(",", ""))
    tot = sum(v for _, v in agg.values())
    out = [f"# launch list summary of {path}", "",
           f"one train step = {hi - lo} launches, {tot / 1e6:.3f} ms of kernel time (ncu: cold-cache, serialised; compare shares)", "",
           "| ms | share | launches | kernel |", "|---:|---:|---:|---|"]
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {v / 1e6:.3f} | {100 * v / tot:.1f}% | {c} | `{k[:120]}` |")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
