"""Per-kernel counts of the Blackwell tensor-path SASS opcodes in libideas_b200.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA load / store), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), UTCATOMSWS / UTCCP (TMEM alloc / copy).  Usage: python scripts/sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ideas_b200", "lib", "libideas_b200.so")
OPS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "UTCATOMSWS", "FFMA", "HMMA", "RED", "ATOM"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
arch = set()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        cur = re.sub(r"\(.*", "", name)
        counts.setdefault(cur, collections.Counter())
        continue
    m = re.search(r"arch = (sm_\w+)", ln)
    if m:
        arch.add(m.group(1))
    if cur:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
print(f"# {os.path.relpath(LIB)}  architectures: {sorted(arch)}")
print(f"{'kernel':70s} " + " ".join(f"{o:>8s}" for o in OPS))
tot = collections.Counter()
for k, c in counts.items():
    if any(c[o] for o in OPS[:7]) or "conv" in k or "blur" in k:
        print(f"{k[:70]:70s} " + " ".join(f"{c[o]:8d}" for o in OPS))
    tot.update(c)
print(f"{'TOTAL (all ' + str(len(counts)) + ' kernels)':70s} " + " ".join(f"{tot[o]:8d}" for o in OPS))
