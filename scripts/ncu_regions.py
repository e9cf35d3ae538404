"""Per-region view of an .ncu-rep source page: splits the SASS of the profiled kernel at branch targets /
barriers into straight-line regions and prints, per region, executed instructions, stall samples and the
dominant stall reasons.  usage: python scripts/ncu_regions.py report.ncu-rep [min_pct]"""
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, data = rows[1], rows[2:]
    ia, isrc, isamp, iaddr = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples"), h.index("Address")
    stall_cols = [(i, c[6:]) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    addrs = [int(r[iaddr], 16) for r in data]
    base = addrs[0]
    targets = set()
    for r in data:
        m = re.search(r"\b(BRA|BSSY)\S*\s+.*?(0x[0-9a-f]+)", r[isrc])
        if m and "BRA" in m.group(1):
            targets.add(int(m.group(2), 16))
    regions, cur = [], []
    for a, r in zip(addrs, data):
        if a in targets and cur:
            regions.append(cur)
            cur = []
        cur.append((a, r))
        if re.search(r"\b(BRA|BAR|EXIT)\b", r[isrc].split(";")[0]) and not r[isrc].strip().startswith("BSSY"):
            regions.append(cur)
            cur = []
    if cur:
        regions.append(cur)
    tot_s = sum(int(r[isamp]) for r in data)
    tot_i = sum(int(r[ia]) for r in data)
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    print("%-16s %5s %7s %7s %6s  %s" % ("range", "lines", "inst%", "samp%", "s/i", "mix / top stalls"))
    for reg in regions:
        s = sum(int(r[isamp]) for _, r in reg)
        i = sum(int(r[ia]) for _, r in reg)
        if 100.0 * s / tot_s < min_pct:
            continue
        st = sorted(((sum(int(r[c]) for _, r in reg), n) for c, n in stall_cols), reverse=True)[:4]
        ops = {}
        for _, r in reg:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
            ops[m.group(2) if m else "?"] = ops.get(m.group(2) if m else "?", 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
        print("%06x-%06x %6d %6.2f%% %6.2f%% %6.2f  %s | %s" % (
            reg[0][0] - base, reg[-1][0] - base, len(reg), 100.0 * i / tot_i, 100.0 * s / tot_s,
            (s / tot_s) / max(i / tot_i, 1e-12),
            " ".join("%s:%d" % kv for kv in top), " ".join("%s:%.0f%%" % (n, 100.0 * v / max(s, 1)) for v, n in st)))


if __name__ == "__main__":
    main()
