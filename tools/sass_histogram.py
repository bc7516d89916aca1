"""SASS opcode histogram of libpvb200.so per kernel family (evidence for the tcgen05 / TMA / TMEM paths):
    python tools/sass_histogram.py > profiles/r02_sass_histogram.md
Runs on the CPU box (cuobjdump only).  Opcodes are counted per kernel; the table keeps the mnemonics
B200_PROFILING.md names as evidence (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, LDGSTS = cp.async, HMMA = mma.sync) plus the ten most frequent others."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorchvideo_b200", "lib", "libpvb200.so")
EVIDENCE = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.2CTA", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCBAR.2CTA", "SYNCS", "LDGSTS",
            "HMMA", "ELECT", "UCGABAR_ARV", "ACQBULK", "LDSM", "REDUX"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("pv::", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    fam = collections.OrderedDict()
    for name, cnt in kernels.items():
        base = re.sub(r"<.*", "", name)
        f = fam.setdefault(base, [0, collections.Counter()])
        f[0] += 1
        f[1].update(cnt)
    print("# SASS opcode histogram of libpvb200.so (sm_100a), per kernel family\n")
    print("`python tools/sass_histogram.py` (cuobjdump -sass, all template instances of a kernel summed).\n")
    print("| kernel | instances | instructions | tcgen05 / TMA / TMEM / mbarrier evidence | most frequent other opcodes |")
    print("|---|---|---|---|---|")
    for base, (n, cnt) in fam.items():
        total = sum(cnt.values())
        ev = []
        for e in EVIDENCE:
            if e.endswith(".2CTA"):
                c = sum(v for k, v in cnt.items() if k.startswith(e.split(".")[0]) and ".2CTA" in k)
            else:
                c = sum(v for k, v in cnt.items() if k.split(".")[0] == e and ".2CTA" not in k)
            if c:
                ev.append("%s x%d" % (e, c))
        rest = collections.Counter()
        for k, v in cnt.items():
            if k.split(".")[0] not in [e.split(".")[0] for e in EVIDENCE]:
                rest[k.split(".")[0]] += v
        top = ", ".join("%s %d" % kv for kv in rest.most_common(8))
        print("| `%s` | %d | %d | %s | %s |" % (base, n, total, ", ".join(ev) or "-", top))


if __name__ == "__main__":
    main()
