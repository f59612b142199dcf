"""Author BASELINE.json config 5 (synthetic quarter core) from the reference's C5G7 2-D example (data file):
an n x n checkerboard of the UO2 / MOX assemblies (lattices 1 / 2) with a reflector ring (lattice 3) on the east
and south sides, west / north reflective -- one 2-D plane (the 2D3D stack repeats it axially).

    python tools/make_quarter_core.py <c5g7_2d.xml> <out.xml> [--n 9] [--max-iter 1]
"""
import argparse
import re

ap = argparse.ArgumentParser()
ap.add_argument("src")
ap.add_argument("dst")
ap.add_argument("--n", type=int, default=9)
ap.add_argument("--max-iter", type=int, default=1)
a = ap.parse_args()
x = open(a.src).read()
n = a.n
rows = []
for j in range(n):
    row = []
    for i in range(n):
        if i == n - 1 or j == n - 1:
            row.append("3")
        else:
            row.append("1" if (i + j) % 2 == 0 else "2")
    rows.append("    " + " ".join(row))
core = (f'<core nx="{n}" ny="{n}" enabled="t"\n    north  = "reflect"\n    south  = "vacuum"\n    east   = "vacuum"\n'
        f'    west   = "reflect"\n    top    = "reflect"\n    bottom = "reflect" >\n' + "\n".join(rows) + "\n</core>")
x, k = re.subn(r"<core .*?</core>", core, x, flags=re.S)
assert k == 1
x, k = re.subn(r'max_iter="\d+"', f'max_iter="{a.max_iter}"', x)
x = x.replace("<case_name>C5G7_2D</case_name>", f"<case_name>QUARTER_CORE_{n}x{n}</case_name>")
open(a.dst, "w").write(x)
print("wrote", a.dst)
