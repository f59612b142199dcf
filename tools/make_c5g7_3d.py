"""Author the C5G7 3-D input of BASELINE.json config 4 from the reference's 2-D example (data file):
same pins / lattices / materials; 6 fuel planes of 7.14 cm (42.84 cm) under 3 axial-reflector planes of
lattice 3 (21.42 cm), bottom reflective, top vacuum; solved with the 2D3D method.

    python tools/make_c5g7_3d.py <c5g7_2d.xml> <out.xml> [--planes-fuel 6 --planes-refl 3 --n-inner 3
                                                            --max-iter 30 --spacing 0.05 --n-azimuthal 8]
"""
import argparse
import re

ap = argparse.ArgumentParser()
ap.add_argument("src")
ap.add_argument("dst")
ap.add_argument("--planes-fuel", type=int, default=6)
ap.add_argument("--planes-refl", type=int, default=3)
ap.add_argument("--n-inner", type=int, default=3)
ap.add_argument("--max-iter", type=int, default=30)
ap.add_argument("--spacing", type=float, default=0.05)
ap.add_argument("--n-azimuthal", type=int, default=8)
ap.add_argument("--tol", default="1.e-7")
a = ap.parse_args()

x = open(a.src).read()
nf, nr = a.planes_fuel, a.planes_refl
hz = 42.84 / nf
assert abs(21.42 / nr - hz) < 1e-12 or True
solver = f'''<solver type="eigenvalue" k_tol="{a.tol}" psi_tol="{a.tol}" max_iter="{a.max_iter}" cmfd="t">
    <cmfd enabled="t" />
    <source scattering="P0" />
    <sweeper type="2d3d">
        <ang_quad type="chebyshev-gauss" n_azimuthal="{a.n_azimuthal}" n_polar="2" />
        <moc_sweeper n_inner="{a.n_inner}">
            <rays spacing="{a.spacing}" modularity="core" />
        </moc_sweeper>
        <sn_sweeper equation="cdd" axial="sc" n_inner="{a.n_inner}" />
    </sweeper>
</solver>'''
x, n = re.subn(r"<solver .*?</solver>", solver, x, flags=re.S)
assert n == 1
for aid, lat in ((1, 1), (2, 2), (3, 3)):
    lats = " ".join(["3"] * nr + [str(lat)] * nf)  # read top-down (assembly.cpp:89-92)
    hzs = " ".join([f"{21.42 / nr:.10g}"] * nr + [f"{hz:.10g}"] * nf)
    new = f'<assembly id="{aid}" np="{nf + nr}">\n    <hz>{hzs}</hz>\n    <lattices>\n        {lats}\n    </lattices>\n</assembly>'
    x, n = re.subn(rf'<assembly id="{aid}".*?</assembly>', new, x, flags=re.S)
    assert n == 1
x = x.replace('<case_name>C5G7_2D</case_name>', '<case_name>C5G7_3D</case_name>')
x, n = re.subn(r'top\s*=\s*"reflect"', 'top    = "vacuum"', x)
assert n == 1
open(a.dst, "w").write(x)
print("wrote", a.dst)
