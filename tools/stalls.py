import csv, sys, subprocess, re, collections
rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; body = []; n = 0
for r in rows:
    if r and r[0] == 'Kernel Name':
        n += 1
        if n > 1: break
    elif r and r[0] == 'Address': hdr = r
    elif hdr and len(r) == len(hdr) and r[0].startswith('0x'): body.append(r)
i_s = hdr.index('Warp Stall Sampling (All Samples)'); i_src = hdr.index('Source'); i_ex = hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[i_s]) for r in body)
print('total samples', tot, 'static instrs', len(body), 'executed warp-instr', sum(int(r[i_ex]) for r in body))
agg = {hdr[i]: sum(int(r[i] or 0) for r in body) for i in stall_cols}
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
# cumulative profile by address decile
cum = 0
for k in range(0, len(body), max(1, len(body) // 20)):
    seg = body[k:k + max(1, len(body) // 20)]
    print('  instr %5d-%5d samples %4d  exec %7d' % (k, k + len(seg), sum(int(r[i_s]) for r in seg), sum(int(r[i_ex]) for r in seg)))
top = sorted(enumerate(body), key=lambda x: -int(x[1][i_s]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]
for idx, r in sorted(top):
    st = {hdr[i][6:]: int(r[i] or 0) for i in stall_cols if int(r[i] or 0) > 0}
    print(idx, r[i_s], r[i_ex], r[i_src][:64], st)
