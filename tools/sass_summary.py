"""Per-kernel SASS evidence for the built library: counts of the mnemonics that prove the Blackwell-specific paths
(UBLKCP = cp.async.bulk / TMA engine, SYNCS = mbarrier, FMNMX3 = 3-input min/max, RED/ATOM = reductions) and the
cubin architecture.  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hiprfish-image-analysis_b200", "libhipr_b200.so")
MNEMONICS = ("UBLKCP", "SYNCS", "FMNMX3", "FMNMX", "VIMNMX", "RED", "ATOM", "DSETP", "MUFU", "LDS", "LDG", "STG", "BAR")


def main():
    elf = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    archs = sorted(set(re.findall(r"sm_\d+a?", elf)))
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", fn).replace("void ", "").replace("hipr::", "")
            per[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and fn:
            op = m.group(1)
            per[fn]["TOTAL"] += 1
            for k in MNEMONICS:
                if op == k or (k in ("RED", "ATOM") and op.startswith(k)):
                    per[fn][k] += 1
    print("libhipr_b200.so: cubin architectures:", ", ".join(archs))
    print("%-72s %7s " % ("kernel", "instrs") + " ".join("%7s" % k for k in MNEMONICS))
    tot = collections.Counter()
    for fn, c in per.items():
        tot.update(c)
        print("%-72s %7d " % (fn[:72], c["TOTAL"]) + " ".join("%7d" % c[k] for k in MNEMONICS))
    print("%-72s %7d " % ("ALL KERNELS", tot["TOTAL"]) + " ".join("%7d" % tot[k] for k in MNEMONICS))


if __name__ == "__main__":
    main()
