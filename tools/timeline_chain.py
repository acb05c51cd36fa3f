#!/usr/bin/env python
"""Main-stream chain analysis of a tools/timeline.py capture: the time each kernel ADDS to the chain (end-to-end on the
busiest stream), idle gaps, the forward / backward / optimizer split.
usage: python tools/timeline_chain.py gpurun_out/timeline_train16f.json [--dump START N]"""
import collections
import json
import sys

sys.path.insert(0, "tools")
from timeline_report import short  # noqa: E402


def main():
    j = json.load(open(sys.argv[1]))
    ev = j["events"]
    streams = collections.Counter(e[3] for e in ev)
    main_id = streams.most_common(1)[0][0]
    rows, prev_end = [], 0.0
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for s, d, n, r in ev:
        if r != main_id:
            continue
        e = s + d
        exp = e - max(prev_end, s) if e > prev_end else 0.0
        gap = max(0.0, s - prev_end)
        k = short(n)
        rows.append((s, d, exp, gap, k))
        agg[k][0] += 1
        agg[k][1] += exp
        agg[k][2] += gap
        prev_end = max(prev_end, e)
    span = prev_end - rows[0][0]
    print(f"main stream {main_id}: {len(rows)} kernels, span {span / 1e3:.3f} ms; other streams: "
          + ", ".join(f"{k}: {v} kernels, busy {sum(e[1] for e in ev if e[3] == k) / 1e3:.2f} ms" for k, v in streams.items() if k != main_id))
    print(f"{'kernel':42s} {'n':>4s} {'exposed ms':>10s} {'avg us':>7s} {'gap-before ms':>13s}")
    for k, v in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
        print(f"{k:42s} {v[0]:4d} {v[1] / 1e3:10.3f} {v[1] / v[0]:7.1f} {v[2] / 1e3:13.3f}")
    marks = {k: i for i, r in enumerate(rows) for k in ("masked_ce", "adamw") if k in r[4]}
    if "masked_ce" in marks and "adamw" in marks:
        ce, ad = rows[marks["masked_ce"]], rows[marks["adamw"]]
        print(f"forward {(ce[0] - rows[0][0]) / 1e3:.3f} ms | backward {(ad[0] - ce[0]) / 1e3:.3f} ms | optimizer {ad[1] / 1e3:.3f} ms")
    if "--dump" in sys.argv:
        a = sys.argv.index("--dump")
        start, n = int(sys.argv[a + 1]), int(sys.argv[a + 2])
        if start < 0:
            start = marks.get("masked_ce", 0) - start
        for r in rows[start:start + n]:
            print(f"  t={r[0]:9.1f} dur={r[1]:6.1f} exposed={r[2]:6.1f} gap={r[3]:5.1f} {r[4]}")


if __name__ == "__main__":
    main()
