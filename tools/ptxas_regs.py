"""Registers / spills of the sweep kernels from the ptxas log: python tools/ptxas_regs.py [substring]"""
import re, sys
pat = sys.argv[1] if len(sys.argv) > 1 else 'rchunk'
txt = open('mocc_b200/csrc/ptxas.log').read()
for b in re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]:
    name = b.split("'")[0]
    if pat not in name:
        continue
    dem = re.findall(r'Li(\d+)E', name)
    regs = re.search(r'Used (\d+) registers', b)
    spill = re.search(r'(\d+) bytes spill stores, (\d+) bytes spill loads', b)
    print(','.join(dem), 'regs', regs.group(1), 'spill', spill.groups() if spill else None)
