"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use, from the in-tree libsvsk.so.

    python tools/sass_summary.py > profiles/r02_libsvsk_sass_summary.txt

UTCHMMA = tcgen05.mma (".2CTA" = cta_group::2), LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG / UTMAREDG =
TMA tensor load / store / reduce, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops
(/opt/skills/guides/B200_PROFILING.md).  Registers and spills come from `cuobjdump -res-usage`."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ensemble_svs_with_interactions_b200", "csrc", "libsvsk.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "UBLKCP", "SYNCS", "MUFU", "HMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    regs = {}
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and fn:
            regs[fn] = (int(m.group(1)), int(m.group(2)))
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
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_total"] += 1
        base = op.split(".")[0]
        if base in KEYS:
            counts[cur][base] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            counts[cur]["UTCHMMA.2CTA"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
    print("# " + " ".join(f"{k:>12}" for k in ["instr", "regs"] + KEYS) + "  kernel")
    tot = collections.Counter()
    for (name, c), dn in zip(counts.items(), demangle):
        short = re.sub(r"\(.*", "", dn).replace("svsk::", "")
        r = regs.get(name, (0, 0))[0]
        print("  " + " ".join(f"{v:>12}" for v in [c["_total"], r] + [c[k] for k in KEYS]) + f"  {short}")
        tot.update(c)
    print("# " + " ".join(f"{v:>12}" for v in [tot["_total"], ""] + [tot[k] for k in KEYS]) + "  TOTAL")


if __name__ == "__main__":
    sys.exit(main())
