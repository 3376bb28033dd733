"""Per-source-line totals (executed warp instructions, stall samples) of one kernel of an ncu report captured with
`--set full --import-source on` from a `-lineinfo` build.
    python tools/srclines.py gpurun_out/x.ncu-rep reset_observe [top_n] [launch_index]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


# the view is a sequence of blocks: "File Path", "Function Name", header, rows...; a new launch restarts with the first file
launch, seen, fname, func, hdr = -1, set(), None, None, None
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        func = r[1]
        key = (fname, func)
        if key in seen or launch < 0:
            launch += 1
            seen = set()
        seen.add(key)
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and launch == which and r[0].isdigit():
        lines.append((fname, int(r[0]), r))
i_s = hdr.index("Warp Stall Sampling (All Samples)")
i_ex = hdr.index("Instructions Executed")
i_src = 1
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot_ex = sum(num(r[i_ex]) for _, _, r in lines)
tot_s = sum(num(r[i_s]) for _, _, r in lines)
print("kernel launch %d: executed warp-instr %d, samples %d" % (which, tot_ex, tot_s))
agg = {}
for f, ln, r in lines:
    k = (f, ln)
    a = agg.setdefault(k, [0, 0, r[i_src], {}])
    a[0] += num(r[i_ex])
    a[1] += num(r[i_s])
    for i in stall_cols:
        if num(r[i]):
            a[3][hdr[i][6:]] = a[3].get(hdr[i][6:], 0) + num(r[i])
byfile = {}
for (f, ln), a in agg.items():
    b = byfile.setdefault(f, [0, 0])
    b[0] += a[0]
    b[1] += a[1]
for f, b in byfile.items():
    print("  %-24s exec %5.1f%%  samples %5.1f%%" % (f, 100.0 * b[0] / max(tot_ex, 1), 100.0 * b[1] / max(tot_s, 1)))
print("top lines by executed instructions:")
for (f, ln), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top_n]:
    top = sorted(a[3].items(), key=lambda x: -x[1])[:3]
    print("  %-18s %4d  exec %5.2f%%  samp %5.2f%%  %-60s %s" % (f, ln, 100.0 * a[0] / max(tot_ex, 1), 100.0 * a[1] / max(tot_s, 1), a[2].strip()[:60], top))
print("top lines by samples:")
for (f, ln), a in sorted(agg.items(), key=lambda x: -x[1][1])[:top_n]:
    top = sorted(a[3].items(), key=lambda x: -x[1])[:3]
    print("  %-18s %4d  exec %5.2f%%  samp %5.2f%%  %-60s %s" % (f, ln, 100.0 * a[0] / max(tot_ex, 1), 100.0 * a[1] / max(tot_s, 1), a[2].strip()[:60], top))
# optional: totals by line range of one file: REGIONS="ti5_post_physics.cu:139-265,318-480"
import os
if os.environ.get("REGIONS"):
    f, spec = os.environ["REGIONS"].split(":")
    for part in spec.split(","):
        lo, hi = map(int, part.split("-"))
        ex = sum(a[0] for (ff, ln), a in agg.items() if ff == f and lo <= ln <= hi)
        sm = sum(a[1] for (ff, ln), a in agg.items() if ff == f and lo <= ln <= hi)
        print("  region %s:%d-%d  exec %5.1f%%  samples %5.1f%%" % (f, lo, hi, 100.0 * ex / max(tot_ex, 1), 100.0 * sm / max(tot_s, 1)))
