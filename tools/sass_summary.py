"""Counts the Blackwell-specific SASS instructions per kernel of libsparkcodec.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma kind::f16, .2CTA = cta_group::2), UTCQMMA (tcgen05.mma kind::f8f6f4: the e5m2 cross terms of the
two-term fp32 mode), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor
loads / stores), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier).  Runs without a GPU.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spark-tts_b200", "libsparkcodec.so")
PATTERNS = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"),
            ("UTCQMMA.2CTA", r"\bUTCQMMA\.2CTA"), ("UTCQMMA", r"\bUTCQMMA\b(?!\.2CTA)"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
            ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("UTCBAR", r"\bUTCBAR"),
            ("SYNCS", r"\bSYNCS"), ("UTMAPF", r"\bUTMAPF|UTMACCTL")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][name] += 1
    names = list(counts)
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, out))
    cols = [n for n, _ in PATTERNS]
    print(f"# SASS instruction counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)")
    print("# " + " ".join(f"{c:>12s}" for c in cols) + "  kernel")
    tot = collections.Counter()
    for k in names:
        c = counts[k]
        tot.update(c)
        short = re.sub(r"sparkcodec::\(anonymous namespace\)::", "", demangle.get(k, k))
        short = re.sub(r"\(.*", "", short)
        print("  " + " ".join(f"{c[x]:12d}" for x in cols) + "  " + short)
    print("# " + " ".join(f"{tot[x]:12d}" for x in cols) + "  TOTAL over %d kernels" % len(names))


if __name__ == "__main__":
    main()
