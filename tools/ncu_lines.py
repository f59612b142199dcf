"""Stall samples of one kernel launch aggregated by CUDA source line.
usage: ncu_lines.py <rep> <kernel regex> <launch skip> <mangled-name substring> [lib.so]
Joins the SASS page of the .ncu-rep (instruction order) with nvdisasm -g line info of the cubin."""
import csv, os, re, subprocess, sys, tempfile, collections
rep, kern, skip, mangled = sys.argv[1:5]
lib = sys.argv[5] if len(sys.argv) > 5 else os.path.join(os.path.dirname(__file__), '..', 'mocc_b200', 'csrc', 'libmocc_b200.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis = []
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith('.cubin')):  # one cubin per translation unit
    d = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    if mangled in d:
        dis = d.splitlines()
        break
lines_of = []  # per instruction: (file, line)
infn = False; cur = ('?', 0)
for ln in dis:
    if ln.startswith('\t.section') or ln.startswith('.section'):
        infn = ('.text.' in ln) and (mangled in ln)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.search(r'/\*[0-9a-f]{4,}\*/', ln) and not ln.strip().startswith('//'):
        lines_of.append(cur)
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kern}', '--launch-skip', skip,
                      '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; data = []
for r in rows:
    if r and r[0] == 'Address':
        if hdr is not None:
            break  # second view repeats the table
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
print('kernel', rows[0][1], '| sass rows', len(data), '| disasm instrs', len(lines_of))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.defaultdict(lambda: collections.Counter())
tot = 0
for i, r in enumerate(data):
    key = lines_of[i] if i < len(lines_of) else ('?', 0)
    n = int(r[ci['# Samples']] or 0); tot += n
    agg[key]['samples'] += n
    agg[key]['inst'] += int(r[ci['Instructions Executed']] or 0)
    for h in stalls:
        agg[key][h[6:]] += int(r[ci[h]] or 0)
src = {}
def text(f, l):
    for d in ('mocc_b200/csrc',):
        p = os.path.join(os.path.dirname(__file__), '..', d, f)
        if os.path.exists(p):
            if p not in src: src[p] = open(p).read().splitlines()
            return src[p][l - 1].strip() if 0 < l <= len(src[p]) else ''
    return ''
print('total samples', tot)
for key, c in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if c['samples'] < tot * float(os.environ.get('NCU_MIN_FRAC', '0.004')):
        continue
    top = ', '.join(f'{k}:{v}' for k, v in c.most_common(6) if k not in ('samples', 'inst') and v > 0)
    print(f"{key[0][:22]:22s}:{key[1]:4d} {c['samples']:5d} {100*c['samples']/tot:5.1f}% inst {c['inst']:8d} | {text(*key)[:60]:60s} | {top}")
