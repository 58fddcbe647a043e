"""Top SASS instructions by stall samples from an `ncu --page source --csv` dump (one kernel)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
iconf = hdr.index('L1 Wavefronts Shared Excessive')
data = rows[2:]
tot = sum(int(r[isamp] or 0) for r in data)
print('total samples', tot)
top = sorted(data, key=lambda r: -int(r[isamp] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for r in top:
    s = int(r[isamp] or 0)
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stalls), reverse=True)[:2]
    print(f'{100*s/tot:5.1f}%  {r[isrc][:90]:90s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]} exc_wf={r[iconf]}')
