"""Static SASS statistics of the built library: per kernel, instruction count and the opcodes of interest
(tensor-core / TMA / bulk-copy / barrier mnemonics from /opt/skills/guides/B200_PROFILING.md), plus the full opcode mix of
kernels selected with --mix.
usage: python scripts/sass_stats.py [LIB.so] [--mix REGEX] [--loop KERNEL_REGEX]  > profiles/sass_opcodes.txt"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "NANOSLEEP", "MUFU", "SHFL",
       "IMAD.WIDE", "FFMA2", "HMMA", "BAR")


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    name, ins = None, []
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            if name:
                yield name, ins
            name, ins = m.group(1), []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and name:
            ins.append(m.group(1).strip())
    if name:
        yield name, ins


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def opcode(text):
    parts = text.split()
    if parts[0].startswith("@"):
        parts = parts[1:]
    return parts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib", nargs="?", default=os.path.join(ROOT, "scrutiny_b200", "libscrutiny_b200.so"))
    ap.add_argument("--mix", default=None, help="regex on the demangled name: print the full opcode mix")
    a = ap.parse_args()
    print("# cuobjdump -sass %s  (static instruction counts per kernel)" % os.path.relpath(a.lib, ROOT))
    for name, ins in kernels(a.lib):
        dn = demangle(name)
        short = re.sub(r"\(.*", "", dn)
        ops = collections.Counter(opcode(i) for i in ins)
        keyed = collections.Counter()
        for op, c in ops.items():
            for k in KEY:
                if op == k or op.startswith(k + ".") or (k == "IMAD.WIDE" and op.startswith("IMAD.WIDE")):
                    keyed[k] += c
        print("%-90s %6d instr  %s" % (short[:90], len(ins), " ".join("%s=%d" % kv for kv in sorted(keyed.items()))))
        if a.mix and re.search(a.mix, dn):
            base = collections.Counter(op.split(".")[0] for op in (opcode(i) for i in ins))
            print("    mix: " + " ".join("%s=%d" % kv for kv in base.most_common()))


if __name__ == "__main__":
    sys.exit(main())
