"""Small driver for ncu captures: a few sweeps of one configuration on a full-size case."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mocc_b200 import Sweeper, load_arrays  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("flat"); ap.add_argument("golden")
ap.add_argument("--batched", type=int, default=1)
ap.add_argument("--boundary", type=int, default=1)
ap.add_argument("--max-polar", type=int, default=2)
ap.add_argument("--tally", type=int, default=0)
ap.add_argument("--n-inner", type=int, default=2)
ap.add_argument("--kernel", type=int, default=-1)
a = ap.parse_args()
flat = load_arrays(a.flat); gold = load_arrays(a.golden)
G = int(flat["n_group"][0]); n_reg = int(flat["n_reg"][0]); bcpg = int(flat["bc_per_group"][0])
xstr = np.stack([gold[f"xs_tr_{g}"] for g in range(G)])
xself = np.stack([gold[f"xs_self_{g}"] for g in range(G)])
kw = {} if a.kernel < 0 else {"kernel": a.kernel}
sw = Sweeper(flat, boundary_update=a.boundary, max_polar=a.max_polar, **kw)
sw.set_xs(0, xstr, xstr_src=xstr, xs_self=xself)
sw.set_source(0, np.full((G, n_reg), 0.1)); sw.set_flux(0, np.ones((G, n_reg)))
for ip in range(sw.n_plane):
    sw.set_boundary(ip, 0, np.full((G, bcpg), 1.0 / (4 * np.pi)))
if a.batched:
    sw.sweep(0, G, n_inner=a.n_inner, tally_mode=a.tally)
else:
    sw.sweep(3, 1, n_inner=a.n_inner, tally_mode=a.tally)
sw.synchronize()
print("done", sw.last_sweep_ms())
