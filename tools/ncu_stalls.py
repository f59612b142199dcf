"""SASS instructions of one kernel launch with the most stall samples, with a few preceding instructions.
usage: ncu_stalls.py <rep> <kernel regex> [launch skip] [min samples] [stall column, e.g. stall_long_sb]"""
import csv, subprocess, sys
rep, kern = sys.argv[1:3]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
mins = int(sys.argv[4]) if len(sys.argv) > 4 else 25
col = sys.argv[5] if len(sys.argv) > 5 else "# Samples"
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kern}', '--launch-skip', skip,
                      '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; data = []
for r in rows:
    if r and r[0] == 'Address':
        if hdr: break
        hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ci['# Samples']] or 0) for r in data)
print('total samples', tot, {h[6:]: sum(int(r[ci[h]] or 0) for r in data) for h in stalls if sum(int(r[ci[h]] or 0) for r in data) > tot * 0.02})
for i, r in enumerate(data):
    if int(r[ci[col]] or 0) >= mins:
        print('----')
        for j in range(max(0, i - 5), i + 1):
            q = data[j]
            st = {h[6:]: int(q[ci[h]] or 0) for h in stalls if int(q[ci[h]] or 0) > 0}
            st = dict(sorted(st.items(), key=lambda x: -x[1])[:3])
            print(f"{j:5d} smp {int(q[ci['# Samples']] or 0):4d} exec {q[ci['Instructions Executed']]:>8s}  {q[ci['Source']][:90]:90s} {st if j == i else ''}")
