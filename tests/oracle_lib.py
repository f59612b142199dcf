"""ctypes access to the CPU oracle (oracle/libmoc_oracle.so). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from mocc_b200.capi import Problem, problem_from_arrays

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libmoc_oracle.so")
_f64p = C.POINTER(C.c_double)

_lib = None


def oracle():
    global _lib
    if _lib is None:
        src = os.path.join(ORACLE_DIR, "moc_oracle.c")
        if (not os.path.exists(ORACLE_LIB)) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "c"], stdout=subprocess.DEVNULL)
        lib = C.CDLL(ORACLE_LIB)
        lib.moc_oracle_exp.argtypes = [_f64p, C.c_int, C.c_double, C.c_double, C.c_double]
        lib.moc_oracle_exp.restype = C.c_double
        lib.moc_oracle_self_scatter.argtypes = [C.c_int] + [_f64p] * 5
        lib.moc_oracle_self_scatter.restype = None
        lib.moc_oracle_fission_source.argtypes = [C.c_int, C.c_int, C.c_double, _f64p, _f64p, _f64p]
        lib.moc_oracle_fission_source.restype = None
        lib.moc_oracle_group_source.argtypes = [C.c_int, C.c_int, C.c_int, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p]
        lib.moc_oracle_group_source.restype = None
        lib.moc_oracle_sweep1g.argtypes = [C.POINTER(Problem), C.c_int, C.c_int, _f64p, _f64p, _f64p, _f64p,
                                           _f64p, _f64p, _f64p]
        lib.moc_oracle_sweep1g_corrections.argtypes = [C.POINTER(Problem), C.c_int] + [_f64p] * 11
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(_f64p)


def oracle_exp(table, v, n=10000, vmin=-10.0, vmax=0.0):
    t = np.ascontiguousarray(table, dtype=np.float64)
    return oracle().moc_oracle_exp(_p(t), n, vmin, vmax, float(v))


def oracle_self_scatter(src, flux, xs_self, xs_tr):
    src, flux, xs_self, xs_tr = (np.ascontiguousarray(x, dtype=np.float64) for x in (src, flux, xs_self, xs_tr))
    q = np.empty_like(src)
    oracle().moc_oracle_self_scatter(src.size, _p(src), _p(flux), _p(xs_self), _p(xs_tr), _p(q))
    return q


def oracle_sweep1g(arrays, xstr, qbar, bc_in, gs_boundary=True, tally_mode=0):
    """Returns (flux_out, bc_after, current, surface_flux); bc_in is [n_plane, bc_per_group]."""
    prob, keep = problem_from_arrays(arrays)
    xstr = np.ascontiguousarray(xstr, dtype=np.float64)
    qbar = np.ascontiguousarray(qbar, dtype=np.float64)
    bc = np.array(bc_in, dtype=np.float64, copy=True).reshape(prob.n_plane, prob.bc_per_group)
    flux = np.zeros(prob.n_reg)
    cur = np.zeros(prob.n_surf)
    sf = np.zeros(prob.n_surf)
    area = np.ascontiguousarray(arrays["surf_area"], dtype=np.float64)
    rc = oracle().moc_oracle_sweep1g(C.byref(prob), int(gs_boundary), tally_mode, _p(xstr), _p(qbar), _p(bc),
                                     _p(flux), _p(cur), _p(sf), _p(area))
    if rc != 0:
        raise RuntimeError("oracle sweep failed")
    return flux, bc, cur, sf


def oracle_sweep1g_corrections(arrays, xstr_split, xstr_true, qbar, sn_xs, bc_in, gs_boundary=True):
    """sweep1g<cmdo::CurrentCorrections>: returns (flux, bc_after, current, surface_flux, alpha, beta)."""
    prob, keep = problem_from_arrays(arrays)
    xs, xt, q, sn = (np.ascontiguousarray(x, dtype=np.float64) for x in (xstr_split, xstr_true, qbar, sn_xs))
    bc = np.array(bc_in, dtype=np.float64, copy=True).reshape(prob.n_plane, prob.bc_per_group)
    n_cell = prob.n_plane * prob.n_cell_plane
    flux, cur, sf = np.zeros(prob.n_reg), np.zeros(prob.n_surf), np.zeros(prob.n_surf)
    alpha = np.full((2 * prob.n_ang, n_cell, 2), np.nan)
    beta = np.full((2 * prob.n_ang, n_cell), np.nan)
    area = np.ascontiguousarray(arrays["surf_area"], dtype=np.float64)
    rc = oracle().moc_oracle_sweep1g_corrections(C.byref(prob), int(gs_boundary), _p(xs), _p(xt), _p(q), _p(sn),
                                                 _p(bc), _p(flux), _p(cur), _p(sf), _p(area), _p(alpha), _p(beta))
    if rc != 0:
        raise RuntimeError(f"oracle corrections sweep failed ({rc})")
    return flux, bc, cur, sf, alpha, beta


def oracle_fission_source(k, xs_nf, flux):
    """TransportSweeper::calc_fission_source: xs_nf [G][n_reg], flux [n_reg][G] -> fs [n_reg]."""
    xs_nf = np.ascontiguousarray(xs_nf, dtype=np.float64)
    flux = np.ascontiguousarray(flux, dtype=np.float64)
    G, n_reg = xs_nf.shape
    assert flux.shape == (n_reg, G)
    fs = np.empty(n_reg)
    oracle().moc_oracle_fission_source(n_reg, G, float(k), _p(xs_nf), _p(flux), _p(fs))
    return fs


def oracle_group_source(group, xs_ch, fs, scat_to, flux, ext=None):
    """Source::initialize_group + fission + in_scatter: xs_ch [n_reg], scat_to [G][n_reg], flux [n_reg][G]."""
    scat_to = np.ascontiguousarray(scat_to, dtype=np.float64)
    flux = np.ascontiguousarray(flux, dtype=np.float64)
    G, n_reg = scat_to.shape
    xs_ch = np.ascontiguousarray(xs_ch, dtype=np.float64)
    fs = np.ascontiguousarray(fs, dtype=np.float64)
    ext = None if ext is None else np.ascontiguousarray(ext, dtype=np.float64)
    src = np.empty(n_reg)
    oracle().moc_oracle_group_source(n_reg, G, int(group), _p(ext) if ext is not None else None, _p(xs_ch), _p(fs),
                                     _p(scat_to), _p(flux), _p(src))
    return src
