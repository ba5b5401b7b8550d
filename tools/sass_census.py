"""Per-kernel SASS opcode census of the shipped library: which kernels use the 5th-generation tensor cores (UTC*MMA), TMEM
loads (LDTM), the TMA engine (UBLKCP / UTMALDG) and how many FFMA they contain.

    python tools/sass_census.py > profiles/sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cvae_gan_b200", "libcvaegan_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "HMMA", "FFMA", "DFMA", "SYNCS", "REDG", "RED", "ATOMG", "LDGSTS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, counts, size = None, collections.OrderedDict(), {}
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            size[cur] = 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            op = m.group(1).split(".")[0]
            counts[cur][op] += 1
            size[cur] += 1
    demangle = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode census of cvae_gan_b200/libcvaegan_b200.so (sm_100a); columns: instructions, then counts of " + ", ".join(OPS))
    print("# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA engine), UTCBAR = tcgen05.commit")
    rows = []
    for (name, c), dn in zip(counts.items(), demangle):
        short = re.sub(r"\(.*", "", dn)
        rows.append((short, size[name], [c.get(o, 0) for o in OPS]))
    rows.sort(key=lambda r: (-r[2][0], -r[1]))
    print(f"{'kernel':78s} {'instr':>8s} " + " ".join(f"{o:>8s}" for o in OPS))
    for short, n, vals in rows:
        print(f"{short[:78]:78s} {n:8d} " + " ".join(f"{v:8d}" for v in vals))


if __name__ == "__main__":
    sys.exit(main())
