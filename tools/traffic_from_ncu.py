"""DRAM traffic per launch of a kernel family from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` pass
over a bench command; merges the result into profiles/r02_kernel_traffic.json (read by bench.py's `roofline.traffic`).
usage: python tools/traffic_from_ncu.py <launches.csv> <workload> <launches per step> <kernel regex> "<note>"
The FIRST step's launches are skipped (allocation / first-touch effects); the next `launches per step` are averaged."""
import csv
import json
import re
import sys
from pathlib import Path


def main():
    path, workload, per_step, rx, note = sys.argv[1], sys.argv[2], int(sys.argv[3]), re.compile(sys.argv[4]), sys.argv[5]
    lines = [l for l in open(path) if not l.startswith("==")]
    by_id = {}
    for r in csv.DictReader(lines):
        if not rx.search(r["Kernel Name"]):
            continue
        m = r["Metric Name"]
        if not m.startswith("dram__bytes"):
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
        by_id.setdefault(int(r["ID"]), 0.0)
        by_id[int(r["ID"])] += v
    ids = sorted(by_id)
    take = ids[per_step:2 * per_step] if len(ids) >= 2 * per_step else ids[-per_step:]
    total = sum(by_id[i] for i in take)
    out = Path(__file__).resolve().parent.parent / "profiles" / "r02_kernel_traffic.json"
    rec = json.loads(out.read_text()) if out.exists() else {}
    rec[workload] = {"dram_bytes_per_launch": total / len(take), "launches": len(take), "dram_bytes_per_step": total,
                     "note": note}
    out.write_text(json.dumps(rec, indent=1) + "\n")
    print(workload, len(ids), "launches seen;", len(take), "averaged;", round(total / len(take) / 1e6, 2), "MB per launch;",
          round(total / 1e9, 3), "GB per step")


if __name__ == "__main__":
    main()
