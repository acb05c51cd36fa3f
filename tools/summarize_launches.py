"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share, average.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [out.md]"""
import collections
import csv
import re
import sys


def short_name(name: str) -> str:
    m = re.search(r"(\w+)<([^>]*)>\(", name)
    if m:
        base, targs = m.group(1), m.group(2)
        if base == "gemm_bf16_kernel":
            a = [t.strip() for t in targs.split(",")]
            return f"gemm_bf16_kernel<BN={a[0]},A_MN={a[1]},B_MN={a[2]}>"
        if base == "sample_logits_kernel":
            return f"sample_logits_kernel<{targs}>"
        return base
    m = re.search(r"(\w+)\(", name)
    return m.group(1) if m else name[:40]


def main():
    path = sys.argv[1]
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
        k = short_name(r["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"source: {path}", f"total {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches "
           "(ncu per-launch times: cold cache, serialised — compare shares, not absolutes)", "",
           "| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {v[0]} | {v[1] / 1e3:.2f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.1f} |")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
