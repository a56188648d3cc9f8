"""Warp-stall breakdown of a kernel by CODE REGION from an ncu report's source page (SASS view).
    ncu -i report.ncu-rep --page source --csv --print-source sass > sass.csv
    python tools/ncu_stall_regions.py sass.csv MARKER[,MARKER...]
Regions are delimited by the first occurrence of each marker substring in the SASS text (e.g. USETMAXREG, LDTM, STS.128);
per region: samples, executed warp instructions, the top stall reasons.  Used for profiles/r02_ncu_coarse_split_stalls.txt."""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    markers = sys.argv[2].split(",") if len(sys.argv) > 2 else []
    cuts = [0]
    for m in markers:
        for i, r in enumerate(data):
            if i > cuts[-1] and m in r[ix["Source"]]:
                cuts.append(i)
                break
    cuts.append(len(data))
    total = sum(int(r[ix["# Samples"]] or 0) for r in data)
    print(f"total samples {total}, {len(data)} SASS instructions")
    for a, b, name in zip(cuts[:-1], cuts[1:], ["(start)"] + markers):
        c = collections.Counter()
        ex = 0
        for r in data[a:b]:
            for h in stalls:
                c[h] += int(r[ix[h]] or 0)
            ex += int(r[ix["Instructions Executed"]] or 0)
        n = sum(c.values())
        top = ", ".join(f"{h[6:]} {v}" for h, v in c.most_common(5) if v)
        print(f"from {name:<16} instr {a:5d}..{b:5d}  samples {n:7d} ({100.0 * n / max(1, total):5.1f} %)  warp-instr {ex:10d}  {top}")
    print("\ntop instructions by samples:")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:12]:
        s = int(r[ix["# Samples"]] or 0)
        st = max(stalls, key=lambda h: int(r[ix[h]] or 0))
        print(f"  {s:7d} ({100.0 * s / max(1, total):4.1f} %)  {r[ix['Source']].strip()[:70]:<70}  {st[6:]}")


if __name__ == "__main__":
    main()
