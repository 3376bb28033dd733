import csv, collections, json, sys
tag = sys.argv[1]
d = json.load(open(f'gpurun_out/bench_{tag}.json'))
print('value %.1f M env-steps/s  ms/step %.4f  e2e %.1f M' % (d['value'] / 1e6, d['ms_per_step'], d['e2e']['value'] / 1e6))
for k, v in d['roofline']['kernels'].items():
    print('  %-30s phase %.1f us  per launch %.2f us' % (k, v['phase_ms'] * 1e3, v['ms_per_launch'] * 1e3))
print('  roofline frac %.3f (%s)  whole-step %.0f GB/s' % (d['roofline']['frac'], d['roofline']['kernel'], d['roofline']['whole_step']['achieved']))
lines = [l for l in open(f'gpurun_out/launches_{tag}.csv') if not l.startswith('==')]
dd = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if 'ti5' in row['Kernel Name']:
        dd[row['Kernel Name'][:40]].append(float(row['Metric Value'].replace(',', '')))
for k, v in dd.items():
    v2 = sorted(v)
    print(f"  ncu {k:42s} n={len(v):4d} median={v2[len(v2)//2]:9.1f} min={min(v):9.1f} max={max(v):9.1f} ns")
