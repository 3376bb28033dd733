"""usage: python tools/sass_hist.py [lib.so] > profiles/rNN_sass_opcodes.txt
Per-kernel SASS opcode histogram of the built library (cuobjdump -sass): instruction count, the 24 most frequent
opcodes, and the opcodes that show how data moves (bulk async copies, mbarrier ops, programmatic-launch control,
128-bit global accesses, named barriers, calls).  No tensor-core opcodes are expected: nothing on this path is a
contraction."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ti5_isaacgym_b200", "libti5step.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
kernels = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(demangle(m.group(1)), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
NOTABLE = ("UBLKCP", "SYNCS", "ACQBULK", "PREEXIT", "LDG.E.128", "STG.E.128", "LDG.E.64", "STG.E.64", "BAR.ARV", "BAR.SYNC",
           "CALL", "CCTL", "ATOMG", "RED", "LDL", "STL", "MUFU", "HMMA", "UTCHMMA", "UTCMMA", "TCGEN")
print(f"# SASS opcode histogram of {os.path.basename(lib)} (sm_100a), cuobjdump -sass")
for name, ops in kernels.items():
    total = sum(ops.values())
    base = collections.Counter()
    for op, n in ops.items():
        base[op.split(".")[0]] += n
    print(f"\n## {name}: {total} instructions")
    print("  top: " + ", ".join(f"{op} {n}" for op, n in base.most_common(24)))
    notable = {k: sum(n for op, n in ops.items() if op.startswith(k)) for k in NOTABLE}
    print("  data movement / control: " + ", ".join(f"{k} {v}" for k, v in notable.items() if v))
