"""Top stall sites of one kernel from an .ncu-rep (SASS view with -lineinfo source mapping)."""
import csv, subprocess, sys, re
rep = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else "."; skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kern}',
                      '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = []
for r in rows:
    if r and r[0] == 'Address':
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci['# Samples']] or 0) for r in data)
print('kernel', rows[0][1] if rows else '?', 'total samples', tot, 'instr rows', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ci[h]] or 0) for r in data) for h in stalls}
print('stall totals:', {k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > tot * 0.01})
srccol = ci.get('Source')
for r in sorted(data, key=lambda r: -int(r[ci['# Samples']] or 0))[:top]:
    st = {h[6:]: int(r[ci[h]] or 0) for h in stalls if int(r[ci[h]] or 0) > 0}
    st = dict(sorted(st.items(), key=lambda x: -x[1])[:3])
    print(f"{int(r[ci['# Samples']]):6d} {100*int(r[ci['# Samples']])/max(tot,1):5.1f}%  {r[srccol][:70]:70s} {st}")
