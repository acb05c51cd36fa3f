"""Static tcgen05 / TMEM / TMA instruction counts per kernel of the built library.
usage: cuobjdump -sass mebt_b200/libmebt_b200.so | python tools/sass_evidence.py > profiles/r01_sass_evidence.md"""
import collections
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "SYNCS", "MUFU", "UTCATOMSWS"]


def main():
    cur, counts = None, collections.defaultdict(collections.Counter)
    for line in sys.stdin:
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            counts[cur][m.group(1).split(".")[0]] += 1
    rows = []
    for fn, c in counts.items():
        if any(c[k] for k in KEYS[:7]):
            name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
            name = name.replace("mebt::(anonymous namespace)::", "").replace("void ", "")
            name = re.sub(r"\(.*", "", name)
            rows.append((name[:72], [c[k] for k in KEYS]))
    rows.sort()
    print("# SASS evidence (`cuobjdump -sass mebt_b200/libmebt_b200.so`, sm_100a): tcgen05 / TMEM / TMA mnemonics per kernel\n")
    print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add,")
    print("LDTM / STTM = tcgen05.ld / tcgen05.st, SYNCS = mbarrier operations, UTCATOMSWS = TMEM allocation.  Static instruction")
    print("counts of each compiled kernel variant (template arguments shown).\n")
    print("| kernel | " + " | ".join(KEYS) + " |")
    print("|---|" + "---:|" * len(KEYS))
    for n, v in rows:
        print(f"| `{n}` | " + " | ".join(str(x) for x in v) + " |")


if __name__ == "__main__":
    main()
