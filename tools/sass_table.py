"""cuobjdump -sass of libsmlvm.so -> per-kernel counts of the mnemonics that show what the kernels are made of
(bulk TMA copies, mbarrier waits, fp64 adds, shared-memory traffic).  Writes a markdown table.
usage: python tools/sass_table.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sparse-marginalization-lvm_b200", "lvmhelpers", "libsmlvm.so")
COLS = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UTCHMMA|UTCQMMA|UTCIMMA", "LDTM|STTM", "HMMA", "DADD", "DMUL", "DFMA", "REDUX", "LDS", "STS",
        "LDG", "STG", "BAR"]


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_table.md")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    arch = set()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for c in COLS:
                if re.match(r"(?:%s)(?:\.|$)" % c, op):
                    cur[c] += 1
    with open(out, "w") as f:
        f.write("# SASS mnemonic counts per kernel of libsmlvm.so (`python tools/sass_table.py`)\n\n")
        f.write("cubin architectures in the library: %s.  Static instruction counts from `cuobjdump -sass`; "
                "`UBLKCP` = `cp.async.bulk` (1-D TMA), `SYNCS` = mbarrier operations, no `UTC*MMA` / `LDTM` / `HMMA` "
                "anywhere: nothing on this path is a dense contraction (BASELINE.json north_star).\n\n" % ", ".join(sorted(arch)))
        f.write("| kernel | SASS instr | " + " | ".join(c.replace("|", "/") for c in COLS) + " |\n")
        f.write("|---|---|" + "---|" * len(COLS) + "\n")
        tot = collections.Counter()
        for name, c in kernels.items():
            f.write("| `%s` | %d | " % (name, c["_total"]) + " | ".join(str(c[k]) if c[k] else "" for k in COLS) + " |\n")
            tot.update(c)
        f.write("| **total (%d kernels)** | %d | " % (len(kernels), tot["_total"]) + " | ".join(str(tot[k]) for k in COLS) + " |\n")
    print(out, len(kernels), "kernels;", {k: tot[k] for k in COLS})


if __name__ == "__main__":
    main()
