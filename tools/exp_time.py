"""Exploratory timing of the C-ABI sweep on a full-size case (not the bench contract).

usage: python tools/exp_time.py <case.mocflat> <case.golden> [--reps N]
Prints one JSON line per configuration. The golden file provides per-FSR cross sections and,
when it holds records, a full-size parity check against the reference's own sweep1g output.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mocc_b200 import Sweeper, load_arrays  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("flat")
    ap.add_argument("golden")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--n-inner", type=int, default=10)
    ap.add_argument("--kernels", default="")
    a = ap.parse_args()
    flat = load_arrays(a.flat)
    gold = load_arrays(a.golden)
    G = int(flat["n_group"][0])
    n_reg = int(flat["n_reg"][0])
    n_plane = int(flat["n_plane"][0])
    S = int(flat["n_seg_reference"][0])
    bcpg = int(flat["bc_per_group"][0])
    xstr = np.stack([gold[f"xs_tr_{g}"] for g in range(G)])
    xself = np.stack([gold[f"xs_self_{g}"] for g in range(G)])
    src = np.full((G, n_reg), 0.1)
    print(json.dumps({"case": a.flat, "S": S, "unique_segments": int(flat["seg_len"].size), "n_reg": n_reg, "G": G,
                      "n_plane": n_plane, "n_ang": int(flat["n_ang"][0]), "n_geom": int(flat["n_geom"][0]),
                      "n_trk": int(flat["trk_bc"].size // 2)}), flush=True)

    # full-size parity against the reference records, if any
    n_rec = int(gold["n_rec"][0]) if "n_rec" in gold else 0
    for variant in [v for v in a.kernels.split(",") if v] or [""]:
        kw = {}
        if variant:
            kw["kernel"] = int(variant)
        for r in range(n_rec):
            p = f"rec{r}_"
            g = int(gold[p + "group"][0])
            mode = int(gold[p + "mode"][0])
            gs = bool(gold["gs_boundary"][0])
            sw = Sweeper(flat, boundary_update=0 if gs else 1, **kw)
            sw.set_xs(g, gold[p + "xstr"])
            sw.set_qbar(g, gold[p + "qbar"])
            bc = gold[p + "bc_in"].reshape(n_plane, bcpg)
            for ip in range(n_plane):
                sw.set_boundary(ip, g, bc[ip])
            sw.sweep(g, 1, n_inner=1, tally_mode=mode, use_qbar=True)
            f = sw.get_flux(g, 1)[0]
            ref = gold[p + "flux_out"]
            err = float(np.max(np.abs(f - ref) / np.abs(ref)))
            bco = np.concatenate([sw.get_boundary(ip, g, 1)[0] for ip in range(n_plane)])
            bref = gold[p + "bc_out"]
            berr = float(np.max(np.abs(bco - bref) / np.maximum(np.abs(bref), 1e-300)))
            out = {"parity_record": r, "variant": variant, "group": g, "mode": mode, "flux_max_rel_err": err,
                   "bc_max_rel_err": berr}
            if mode == 1:
                cur, sf = sw.get_coarse(g)
                area = flat["surf_area"]
                m = gold[p + "current"] != 0
                out["current_max_rel_err"] = float(np.max(np.abs(cur[m] / area[m] - gold[p + "current"][m]) /
                                                         np.abs(gold[p + "current"][m])))
                m = gold[p + "surface_flux"] != 0
                out["surface_flux_max_rel_err"] = float(np.max(np.abs(sf[m] / area[m] - gold[p + "surface_flux"][m]) /
                                                              np.abs(gold[p + "surface_flux"][m])))
            print(json.dumps(out), flush=True)
            sw.close()

    configs = []
    for variant in [v for v in a.kernels.split(",") if v] or [""]:
        for bu in (0, 1):
            for mp in (1, 2):
                for batched in (False, True):
                    for tally in (0, 1):
                        configs.append((variant, bu, mp, batched, tally))
    for variant, bu, mp, batched, tally in configs:
        kw = {}
        if variant:
            kw["kernel"] = int(variant)
        sw = Sweeper(flat, boundary_update=bu, max_polar=mp, **kw)
        sw.set_xs(0, xstr, xstr_src=xstr, xs_self=xself)
        sw.set_source(0, src)
        sw.set_flux(0, np.ones((G, n_reg)))
        bc = np.full((G, bcpg), 1.0 / (4 * np.pi))
        for ip in range(n_plane):
            sw.set_boundary(ip, 0, bc)

        def step():
            if batched:
                sw.sweep(0, G, n_inner=a.n_inner, tally_mode=tally)
            else:
                for g in range(G):
                    sw.sweep(g, 1, n_inner=a.n_inner, tally_mode=tally)
        step()
        sw.synchronize()
        best = 1e30
        for _ in range(a.reps):
            t0 = time.perf_counter()
            step()
            sw.synchronize()
            best = min(best, time.perf_counter() - t0)
        upd = 2.0 * S * G * a.n_inner
        st = sw.stats()
        print(json.dumps({"variant": variant, "boundary": "gs" if bu == 0 else "jacobi", "max_polar": mp,
                          "batched": batched, "tally": tally, "ms_per_step": best * 1e3,
                          "updates_per_s": upd / best, "last_sweep_ms": sw.last_sweep_ms(),
                          "frac_hbm_6.1B": upd / best * 6.10 / 6547.8e9,
                          "device_bytes": st["device_bytes"]}), flush=True)
        sw.close()


if __name__ == "__main__":
    main()
