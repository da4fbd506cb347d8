"""Per-kernel SASS opcode summary of the built library (evidence that the hot kernels use tcgen05 / TMEM / bulk-TMA):

    python tools/sass_summary.py > profiles/r02_sass_summary.txt

Counts, per kernel of yoho_b200/libyoho_b200.so (cuobjdump -sass), the mnemonics B200_PROFILING.md lists: UTCHMMA / UTCQMMA
(tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk), UTMALDG / UTMASTG (tensor-map TMA),
LDGSTS (cp.async), SYNCS (mbarrier), HMMA (legacy warp MMA), plus FFMA / DFMA as the SIMT pipes' markers."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "yoho_b200", "libyoho_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "HMMA", "FFMA", "DFMA", "ATOM", "RED", "SHFL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*$", "", name)
            cur = per.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for o in OPS:
                if op.startswith(o):
                    cur[o] += 1
    tot = collections.Counter()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instructions per kernel (sm_100a)")
    print("%-58s %7s " % ("kernel", "instrs") + " ".join("%7s" % o for o in OPS))
    for name, c in per.items():
        print("%-58s %7d " % (name[:58], c["_total"]) + " ".join("%7d" % c[o] for o in OPS))
        tot.update(c)
    print("%-58s %7d " % ("TOTAL", tot["_total"]) + " ".join("%7d" % tot[o] for o in OPS))


if __name__ == "__main__":
    sys.exit(main())
