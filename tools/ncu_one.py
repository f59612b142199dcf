"""ncu driver: a few per-group inner sweeps of C5G7-2D (bench workload) with one kernel selection."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mocc_b200 import Sweeper, load_arrays  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kernel", type=int, default=0)
ap.add_argument("--group", type=int, default=3)
ap.add_argument("--n-inner", type=int, default=2)
ap.add_argument("--tally", type=int, default=0)
ap.add_argument("--boundary", type=int, default=0)
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
arr = load_arrays(bench.workload_files())
G = int(arr["n_group"][0]); n_reg = int(arr["n_reg"][0]); bcpg = int(arr["bc_per_group"][0])
sw = Sweeper(arr, boundary_update=a.boundary, kernel=a.kernel)
sw.set_xs(0, arr["xs_tr"], xstr_src=arr["xs_tr"], xs_self=arr["xs_self"])
sw.set_source(0, bench.synthetic_source(arr, G, n_reg)); sw.set_flux(0, np.ones((G, n_reg)))
for ip in range(sw.n_plane):
    sw.set_boundary(ip, 0, np.full((G, bcpg), 1.0 / (4 * np.pi)))
if a.tally == 2:  # 2D3D correction factors need the homogenised Sn cross sections (any positive values do here)
    ncell = int(arr["n_plane"][0]) * int(arr["n_cell_plane"][0])
    sw.set_sn_xs(0, np.full((G, ncell), 0.5))
for _ in range(a.reps):
    sw.sweep(a.group, 1, n_inner=a.n_inner, tally_mode=a.tally)
    sw.synchronize()
    print("last inner sweep ms", sw.last_sweep_ms())
