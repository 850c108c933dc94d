"""Developer tool: which kernels of libabl_b200.so use tcgen05 / TMEM / TMA, from the SASS of the built library
(cuobjdump -sass).  Prints, per kernel, the counts of the mnemonics that prove it (UTC*MMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store, SYNCS = mbarrier) and
one sample line of each.  python tools/sass_summary.py > profiles/r2_sass_tcgen05_tma_summary.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ab_line_classifier_b200", "libabl_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTCCP", "UTMALDG", "UTMASTG",
        "UTMAPF", "UBLKCP", "SYNCS", "ATOM", "RED."]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fn, counts, samples = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            counts.setdefault(fn, collections.Counter())
            continue
        if fn is None or "/*" not in line:
            continue
        ins = line.split("*/", 1)[-1].strip()
        for k in KEYS:
            if re.search(r"\b" + re.escape(k), ins):
                counts[fn][k] += 1
                samples.setdefault((fn, k), ins.split("/*")[0].strip().rstrip(";"))
    print(f"# SASS of {os.path.basename(LIB)} (sm_100a), mnemonic counts per kernel; float atomics: none expected "
          "(ATOM/RED counts below are integer tickets and metric counters)")
    tot = collections.Counter()
    for fn, c in counts.items():
        if not c:
            continue
        tot.update(c)
        print(f"\n{fn}")
        print("   " + "  ".join(f"{k}:{v}" for k, v in c.items()))
        for k in c:
            print(f"      e.g. {samples[(fn, k)]}")
    print("\n# totals: " + "  ".join(f"{k}:{v}" for k, v in tot.items()))


if __name__ == "__main__":
    sys.exit(main())
