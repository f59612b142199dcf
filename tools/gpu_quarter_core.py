"""BASELINE.json config 5 at plane level: a 9x9-assembly quarter core (233 M reference segments per group sweep,
1.23 M FSRs, tracks of up to 2688 segments): full-size parity against the C oracle and per-inner sweep time."""
import json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mocc_b200 import Sweeper, load_arrays  # noqa: E402
from oracle_lib import oracle_sweep1g  # noqa: E402

W = "/tmp/qc"; os.makedirs(W, exist_ok=True)
inputs = os.path.join(ROOT, "mocc_b200", "bin", "inputs")
subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_quarter_core.py"), os.path.join(inputs, "c5g7_2d.xml"),
                       os.path.join(W, "qc9.xml")], stdout=subprocess.DEVNULL)
subprocess.check_call(["cp", os.path.join(inputs, "c5g7.xsl"), W])
t0 = time.time()
if not os.path.exists(os.path.join(W, "qc9.mocflat")):
    subprocess.check_call([os.path.join(ROOT, "mocc_b200", "bin", "mocc_flatten"), "qc9.xml", "qc9.mocflat", "--xs"], cwd=W,
                          stdout=subprocess.DEVNULL)
t_flat = time.time() - t0
flat = load_arrays(os.path.join(W, "qc9.mocflat"))
G, n_reg, bcpg = (int(flat[k][0]) for k in ("n_group", "n_reg", "bc_per_group"))
S = int(flat["n_seg_reference"][0])
rng = np.random.default_rng(9)
out = {"workload": "quarter core 9x9 assemblies (C5G7 lattices), one plane", "segments": S, "n_reg": n_reg,
       "flatten_s": round(t_flat, 1)}
for kernel in [int(k) for k in os.environ.get("QC_KERNELS", "0,3").split(",")]:
    t0 = time.time()
    sw = Sweeper(flat, boundary_update=0, kernel=kernel)
    out[f"create_s_k{kernel}"] = round(time.time() - t0, 1)
    sw.set_xs(0, flat["xs_tr"], xstr_src=flat["xs_tr"], xs_self=flat["xs_self"])
    errs = []
    for g, tally in ((3, 1), (5, 0)):
        q = rng.uniform(0.05, 1.0, n_reg); bc = rng.uniform(0, 0.3, (1, bcpg))
        sw.set_qbar(g, q); sw.set_boundary(0, g, bc[0])
        sw.sweep(g, 1, n_inner=1, tally_mode=tally, use_qbar=True)
        f_o, bc_o, cur_o, sf_o = oracle_sweep1g(flat, flat["xs_tr"][g], q, bc, gs_boundary=True, tally_mode=tally)
        f = sw.get_flux(g, 1)[0]; b = sw.get_boundary(0, g, 1)[0]
        errs.append(float(np.max(np.abs(f - f_o) / np.abs(f_o))))
        errs.append(float(np.max(np.abs(b - bc_o[0]) / np.maximum(np.abs(bc_o[0]), 1e-30))))
    out[f"max_rel_err_vs_oracle_k{kernel}"] = max(errs)
    sw.set_source(0, np.full((G, n_reg), 0.1)); sw.set_flux(0, np.ones((G, n_reg)))
    sw.set_timing(True)
    for _ in range(2):
        sw.sweep(3, 1, n_inner=10, tally_mode=1)
    ms, n = sw.get_timing()
    out[f"ms_per_inner_k{kernel}"] = ms / n
    out[f"updates_per_s_k{kernel}"] = 2.0 * S / (ms / n * 1e-3)
    out[f"kernel_in_use_k{kernel}"] = int(sw.stats()["kernel"])
    sw.close()
print(json.dumps(out))
