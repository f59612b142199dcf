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
ap.add_argument("--lattice-n", type=int, default=17, help="crop every 17x17 lattice to its central NxN pins (reduced model)")
ap.add_argument("--coarse-pins", action="store_true", help="2+1 rings x 4 sectors per pin instead of 5+2 x 8")
ap.add_argument("--sweeper-attrs", default="", help='extra attributes of <sweeper type="{a.sweeper_type}" {a.sweeper_attrs}>, e.g. \'relax="0.5" cycle="v"\'')
ap.add_argument("--moc-attrs", default="", help="extra attributes of <moc_sweeper>")
ap.add_argument("--sn-attrs", default="", help="extra attributes of <sn_sweeper>")
ap.add_argument("--sn-inner", type=int, default=0)
ap.add_argument("--sn-axial", default="sc")
ap.add_argument("--cmfd-attrs", default="")
ap.add_argument("--top", default="vacuum")
ap.add_argument("--fuel-height", type=float, default=42.84)
ap.add_argument("--refl-height", type=float, default=21.42)
ap.add_argument("--sweeper-type", default="2d3d")
a = ap.parse_args()

x = open(a.src).read()
nf, nr = a.planes_fuel, a.planes_refl
hz = a.fuel_height / nf
solver = f'''<solver type="eigenvalue" k_tol="{a.tol}" psi_tol="{a.tol}" max_iter="{a.max_iter}" cmfd="t">
    <cmfd enabled="t" {a.cmfd_attrs} />
    <source scattering="P0" />
    <sweeper type="{a.sweeper_type}" {a.sweeper_attrs}>
        <ang_quad type="chebyshev-gauss" n_azimuthal="{a.n_azimuthal}" n_polar="2" />
        <moc_sweeper n_inner="{a.n_inner}" {a.moc_attrs}>
            <rays spacing="{a.spacing}" modularity="core" />
        </moc_sweeper>
        <sn_sweeper equation="cdd" axial="{a.sn_axial}" n_inner="{a.sn_inner or a.n_inner}" {a.sn_attrs} />
    </sweeper>
</solver>'''
x, n = re.subn(r"<solver .*?</solver>", solver, x, flags=re.S)
assert n == 1
for aid, lat in ((1, 1), (2, 2), (3, 3)):
    lats = " ".join(["3"] * nr + [str(lat)] * nf)  # read top-down (assembly.cpp:89-92)
    hzs = " ".join([f"{a.refl_height / nr:.10g}"] * nr + [f"{hz:.10g}"] * nf)
    new = f'<assembly id="{aid}" np="{nf + nr}">\n    <hz>{hzs}</hz>\n    <lattices>\n        {lats}\n    </lattices>\n</assembly>'
    x, n = re.subn(rf'<assembly id="{aid}".*?</assembly>', new, x, flags=re.S)
    assert n == 1
x = x.replace('<case_name>C5G7_2D</case_name>', '<case_name>C5G7_3D</case_name>')
x, n = re.subn(r'top\s*=\s*"reflect"', f'top    = "{a.top}"', x)
assert n == 1
if a.lattice_n != 17:
    lo = (17 - a.lattice_n) // 2
    def crop(m):
        rows = [r.split() for r in m.group(2).strip().splitlines()]
        rows = [r[lo:lo + a.lattice_n] for r in rows[lo:lo + a.lattice_n]]
        body = "\n".join("        " + " ".join(r) for r in rows)
        return f'<lattice id="{m.group(1)}" nx="{a.lattice_n}" ny="{a.lattice_n}">\n{body}\n</lattice>'
    x, n = re.subn(r'<lattice id="(\d+)" nx="17" ny="17">(.*?)</lattice>', crop, x, flags=re.S)
    assert n == 3
if a.coarse_pins:
    x = x.replace("<sub_radii>5 2</sub_radii>", "<sub_radii>2 1</sub_radii>").replace("<sub_azi>8</sub_azi>", "<sub_azi>4</sub_azi>")
    x = x.replace("<sub_x>3</sub_x>", "<sub_x>2</sub_x>").replace("<sub_y>3</sub_y>", "<sub_y>2</sub_y>")
    x = re.sub(r'(<pin id="6" mesh="2">).*?(</pin>)', r'\1\n 6 6\n 6 6\n\2', x, flags=re.S)
open(a.dst, "w").write(x)
print("wrote", a.dst)
