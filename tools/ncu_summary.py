"""Compact per-kernel summary of an ncu --set full report (or of its exported raw CSV page).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep|prof_raw.csv [out.md]"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%")]


def main():
    rep = sys.argv[1]
    if rep.endswith(".csv"):          # already exported on the GPU box: ncu -i x.ncu-rep --page raw --csv > x.csv
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = [f"source: {rep} (ncu --set full --clock-control none)", "",
           "| kernel | grid | " + " | ".join(s for _, s in WANT) + " |", "|---|---|" + "---:|" * len(WANT)]
    for r in rows[2:]:
        name = r[col["Kernel Name"]].replace("unnamed>::", "").replace("void ", "")
        name = name.split("(")[0][:60]
        cells = []
        for m, _ in WANT:
            if m in col:
                v, u = r[col[m]], units[col[m]]
                try:
                    cells.append(f"{float(v):.4g} {u}".strip())
                except ValueError:
                    cells.append(v)
            else:
                cells.append("-")
        out.append(f"| {name} | {r[col['Grid Size']]} | " + " | ".join(cells) + " |")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
