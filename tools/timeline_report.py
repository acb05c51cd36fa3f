#!/usr/bin/env python
"""Tables from a tools/timeline.py capture: per-kernel totals, per-stream busy time, idle gaps of the whole device.
usage: python tools/timeline_report.py gpurun_out/timeline_train16f.json [out.md]"""
import collections
import json
import re
import sys


def short(name):
    m = re.match(r"(?:void )?(?:mebt::)?(?:\(anonymous namespace\)::)?(\w+)(<[^(]*>)?", name)
    base = m.group(1) if m else name[:40]
    targs = (m.group(2) or "") if m else ""
    if base == "gemm_bf16_kernel":
        a = [t.strip() for t in targs.strip("<>").split(",")]
        flag = lambda v: {"false": "0", "true": "1"}.get(v, v)
        return f"gemm<BN={a[0]},A_MN={flag(a[1])},B_MN={flag(a[2])}" + (",pair" if len(a) > 4 and flag(a[4]) == "1" else "") + ">"
    return base


def main():
    j = json.load(open(sys.argv[1]))
    ev = j["events"]
    t_end = max(s + d for s, d, _, _ in ev)
    agg = collections.defaultdict(lambda: [0, 0.0])
    streams = collections.defaultdict(float)
    for s, d, n, r in ev:
        k = short(n)
        agg[k][0] += 1
        agg[k][1] += d
        streams[r] += d
    # device-level busy/idle: union of all intervals
    iv = sorted((s, s + d) for s, d, _, _ in ev)
    busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
    gaps = []
    for s, e in iv[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append(s - cur_e)
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    out = [f"source: {sys.argv[1]} ({j['workload']}, {j['steps']} step(s); CUPTI activity records through torch.profiler)",
           f"span {t_end / 1e3:.3f} ms, {len(ev)} device activities, device busy (union over streams) {busy / 1e3:.3f} ms, "
           f"idle gaps {sum(gaps) / 1e3:.3f} ms in {len(gaps)} gaps (median {sorted(gaps)[len(gaps) // 2] if gaps else 0:.1f} us)", "",
           "| stream | busy ms |", "|---|---:|"]
    for r, v in sorted(streams.items(), key=lambda kv: -kv[1]):
        out.append(f"| {r} | {v / 1e3:.3f} |")
    out += ["", "| kernel | launches | total ms | share of span | avg us |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {v[0]} | {v[1] / 1e3:.3f} | {100 * v[1] / t_end:.1f}% | {v[1] / v[0]:.1f} |")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
