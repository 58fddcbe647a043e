#!/usr/bin/env python
"""Turn the two ncu captures of scripts/ncu_session.sh into the tracked summaries under profiles/:
     <tag>_ncu_launches.csv        the launch list (copied)
     <tag>_ncu_launch_shares.txt   share of the step per kernel
     <tag>_ncu_summary.txt         per-launch DRAM bytes / %, L2 %, tensor-pipe %, issue-slot % (from --set full)
     <tag>_ncu_source_fused.txt   source page: stall reasons and hottest SASS lines of the fused layer kernel
     kernel_traffic.json           measured DRAM bytes per launch and subdomain per kernel class (bench.py's roofline.traffic)
   Usage: python scripts/ncu_profiles.py <tag> <domains in the capture>"""
import csv, io, json, os, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, domains = sys.argv[1], int(sys.argv[2])
O, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
CLASS = {'k_tc_relax': 'relax', 'k_tc_update': 'update', 'k_tc_prop': 'prop', 'k_tc_input_embed': 'input', 'k_tc_input_update': 'input',
         'k_tc_fused': 'layer', 'k_tc_fused_input': 'input'}

# ---- launch list ----
lines = [l for l in open(os.path.join(O, f'{tag}_launches.csv')) if not l.startswith('==')]
open(os.path.join(P, f'{tag}_ncu_launches.csv'), 'w').writelines(lines)
rows = list(csv.DictReader(io.StringIO(''.join(lines))))
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = r['Kernel Name'].split('(')[0].split('::')[-1]
    v = float(r['Metric Value'].replace(',', ''))
    v = v / 1e3 if r['Metric Unit'] in ('ns', 'nsecond') else v
    tot[name] += v
    cnt[name] += 1
allus = sum(tot.values())
with open(os.path.join(P, f'{tag}_ncu_launch_shares.txt'), 'w') as f:
    f.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none: one step of bench.py --domains {domains} --chunk {domains} '
            f'(launch list: {tag}_ncu_launches.csv)\n')
    f.write(f"{'kernel':28s} {'launches':>8s} {'total_us':>10s} {'share':>7s}\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write(f'{k:28s} {cnt[k]:8d} {v:10.1f} {100 * v / allus:6.1f}%\n')
    f.write(f"{'all':28s} {sum(cnt.values()):8d} {allus:10.1f}\n")

# ---- full capture ----
raw = subprocess.run(['ncu', '-i', os.path.join(O, f'{tag}_prof.ncu-rep'), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
summ = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_raw_summary.py')], input=raw, capture_output=True, text=True).stdout
with open(os.path.join(P, f'{tag}_ncu_summary.txt'), 'w') as f:
    f.write(f'# ncu --set full --clock-control none, bench.py --domains {domains} --chunk {domains} --steps 1 --warmup 1 (second step)\n')
    f.write(summ)
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]
col = {k: hdr.index(k) for k in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum')}
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
sb = scale[units[col['dram__bytes_read.sum']]]
tu = {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3}[units[col['gpu__time_duration.sum']]]
agg = defaultdict(lambda: [0, 0.0, 0.0])
for r in rr[2:]:
    name = r[col['Kernel Name']].split('(')[0].split('::')[-1]
    c = CLASS.get(name)
    if c is None:
        continue
    a = agg[c]
    a[0] += 1
    a[1] += (float(r[col['dram__bytes_read.sum']].replace(',', '')) + float(r[col['dram__bytes_write.sum']].replace(',', ''))) * sb
    a[2] += float(r[col['gpu__time_duration.sum']].replace(',', '')) * tu
kt = {c: {'launches_per_step': a[0], 'dram_bytes_per_launch_per_subdomain': a[1] / a[0] / domains,
          'dram_bytes_per_step_per_subdomain': a[1] / domains, 'ncu_time_us_per_step': a[2]} for c, a in agg.items()}
kt['_source'] = (f'ncu --set full, profiles/{tag}_ncu_summary.txt: bench.py --domains {domains} --chunk {domains} --steps 1 --warmup 1, second step '
                 '(dram__bytes_read.sum + dram__bytes_write.sum)')
json.dump(kt, open(os.path.join(P, 'kernel_traffic.json'), 'w'), indent=1)
# ---- source page of the two big kernels: stall reasons and hottest SASS lines ----
for kern, short in (('k_tc_fused', 'fused'),):
    src = subprocess.run(['ncu', '-i', os.path.join(O, f'{tag}_prof.ncu-rep'), '--page', 'source', '--csv', '--kernel-name', f'regex:{kern}$',
                          '--launch-skip', '0', '--launch-count', '1'], capture_output=True, text=True).stdout
    tmp = os.path.join(O, f'{tag}_src_{short}.csv')
    open(tmp, 'w').write(src)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_src_summary.py'), tmp, '20'], capture_output=True, text=True).stdout
    with open(os.path.join(P, f'{tag}_ncu_source_{short}.txt'), 'w') as f:
        f.write(f'# ncu --page source of {tag}_prof.ncu-rep, first launch of {kern}: stall reasons and hottest SASS lines\n')
        f.write(out)
print(open(os.path.join(P, f'{tag}_ncu_launch_shares.txt')).read())
print(json.dumps({k: v for k, v in kt.items() if k != '_source'}, indent=1))
