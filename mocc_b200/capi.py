"""ctypes binding of the C ABI declared in ``include/mocc_b200.h``.

This is the FFI a Python host would use; the C++ plugin in ``mocc_b200/host/``
calls the very same entry points.  There is NO CPU fallback: if the CUDA
library is missing or a call fails, a ``RuntimeError`` is raised.
"""
import ctypes as C
import os

import numpy as np

from .flatfile import scalar

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libmocc_b200.so")

MAX_POLAR = 4
TALLY_NONE, TALLY_CURRENT, TALLY_CORRECTIONS = 0, 1, 2
BOUNDARY_GS, BOUNDARY_JACOBI = 0, 1
EXP_TABLE, EXP_FACTORED = 0, 1
KERNEL_AUTO, KERNEL_ITEM, KERNEL_TRACK, KERNEL_CACHED, KERNEL_CHUNK, KERNEL_RCHUNK = 0, 1, 2, 3, 4, 5

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u32p = C.POINTER(C.c_uint32)
_f64p = C.POINTER(C.c_double)

# (field name, ctype) in the exact order of struct mocb200_problem
_SCALARS = [
    ("n_group", C.c_int32), ("n_reg", C.c_int32), ("n_plane", C.c_int32), ("n_unique", C.c_int32),
    ("ndir_oct", C.c_int32), ("n_ang", C.c_int32), ("n_geom", C.c_int32), ("bc_per_group", C.c_int32),
    ("n_surf", C.c_int32), ("n_cell", C.c_int32), ("n_surf_plane", C.c_int32), ("n_cell_plane", C.c_int32),
    ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("exp_n", C.c_int32),
    ("exp_min", C.c_double), ("exp_max", C.c_double),
    ("n_trk", C.c_int64), ("n_seg", C.c_int64), ("n_cm", C.c_int64),
]
_ARRAYS = [
    ("ang_geom", _i32p), ("ang_rsintheta", _f64p),
    ("wt_v_st", _f64p), ("cur_wx", _f64p), ("cur_wy", _f64p), ("flx_wx", _f64p), ("flx_wy", _f64p),
    ("bc_offset", _i32p), ("bc_size_x", _i32p), ("bc_size_y", _i32p), ("bc_dst_off", _i32p),
    ("bc_dst_kind", _i32p),
    ("geom_trk_begin", _i64p), ("trk_seg_begin", _i64p), ("trk_bc", _i32p), ("trk_cm_begin", _i64p),
    ("trk_cm_start", _i32p), ("seg_len", _f64p), ("seg_fsr", _i32p), ("cm_data", _u32p),
    ("plane_unique", _i32p), ("plane_first_reg", _i32p), ("plane_cell_offset", _i32p),
    ("plane_surf_offset", _i32p),
    ("coarse_surf", _i32p), ("coarse_nbr", _i32p),
    ("vol", _f64p), ("exp_table", _f64p),
    ("ang_area_x", _f64p), ("ang_area_y", _f64p), ("ang_ox", _f64p), ("cell_dx", _f64p), ("cell_dy", _f64p),
]
_OPTIONAL = {"ang_area_x", "ang_area_y", "ang_ox", "cell_dx", "cell_dy"}
_NP = {_i32p: np.int32, _i64p: np.int64, _u32p: np.uint32, _f64p: np.float64}


class Problem(C.Structure):
    _fields_ = _SCALARS + _ARRAYS


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("boundary_update", C.c_int32), ("exp_mode", C.c_int32),
                ("max_polar", C.c_int32), ("block_threads", C.c_int32), ("plane_begin", C.c_int32),
                ("plane_end", C.c_int32), ("kernel", C.c_int32), ("chunk_cap", C.c_int32), ("cache_groups", C.c_int32), ("persistent", C.c_int32),
                ("family_begin", C.c_int32), ("family_end", C.c_int32), ("reserved", C.c_int32 * 3)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("sweep_launches", C.c_int64),
                ("segments_per_sweep", C.c_int64), ("unique_segments", C.c_int64),
                ("device_bytes", C.c_int64), ("items", C.c_int64 * 2), ("kernel", C.c_int64),
                ("swept_segments", C.c_int64)]


def problem_from_arrays(arrays):
    """Build a ``Problem`` struct viewing the numpy arrays of a flattened problem.

    Returns (struct, keepalive list); the arrays must outlive the struct's use.
    """
    p = Problem()
    keep = []
    for name, ct in _SCALARS:
        if name == "n_trk":
            v = arrays["trk_bc"].size // 2
        elif name == "n_seg":
            v = arrays["seg_len"].size
        elif name == "n_cm":
            v = arrays["cm_data"].size
        else:
            v = scalar(arrays, name)
        setattr(p, name, v)
    for name, ct in _ARRAYS:
        if name in _OPTIONAL and name not in arrays:
            continue  # stays NULL: no 2D3D correction tally for this problem
        a = np.ascontiguousarray(arrays[name], dtype=_NP[ct]).reshape(-1)
        if a.size == 0:
            a = np.zeros(1, dtype=_NP[ct])
        keep.append(a)
        setattr(p, name, a.ctypes.data_as(ct))
    return p, keep


_lib = None


def load_library(path=None):
    """dlopen the CUDA library; raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"mocc_b200: CUDA library not found at {path}; build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    lib = C.CDLL(path)
    H = C.c_void_p
    lib.mocb200_create.argtypes = [C.POINTER(Problem), C.POINTER(Options), C.POINTER(H)]
    lib.mocb200_destroy.argtypes = [H]
    lib.mocb200_last_error.argtypes = [H]
    lib.mocb200_last_error.restype = C.c_char_p
    lib.mocb200_set_stream.argtypes = [H, C.c_void_p]
    lib.mocb200_synchronize.argtypes = [H]
    lib.mocb200_set_xs.argtypes = [H, C.c_int, C.c_int, _f64p, _f64p, _f64p]
    for fn in ("mocb200_set_source", "mocb200_set_flux", "mocb200_get_flux", "mocb200_set_qbar"):
        getattr(lib, fn).argtypes = [H, C.c_int, C.c_int, _f64p]
    lib.mocb200_set_boundary.argtypes = [H, C.c_int, C.c_int, C.c_int, _f64p]
    lib.mocb200_get_boundary.argtypes = [H, C.c_int, C.c_int, C.c_int, _f64p]
    lib.mocb200_sweep.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.mocb200_get_coarse.argtypes = [H, C.c_int, _f64p, _f64p]
    lib.mocb200_set_sn_xs.argtypes = [H, C.c_int, C.c_int, _f64p]
    lib.mocb200_get_corrections.argtypes = [H, C.c_int, _f64p, _f64p]
    lib.mocb200_set_sweep_inputs.argtypes = [H, C.c_int, _f64p, _f64p, C.POINTER(_f64p)]
    lib.mocb200_get_sweep_results.argtypes = [H, C.c_int, _f64p, C.POINTER(_f64p), _f64p, _f64p]
    lib.mocb200_pack_results_device.argtypes = [H, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.mocb200_get_stats.argtypes = [H, C.POINTER(Stats)]
    lib.mocb200_last_sweep_ms.argtypes = [H, _f64p]
    lib.mocb200_set_timing.argtypes = [H, C.c_int]
    lib.mocb200_get_timing.argtypes = [H, _f64p, C.POINTER(C.c_int64)]
    i32p = C.POINTER(C.c_int32)
    lib.mocb200_get_source.argtypes = [H, C.c_int, C.c_int, _f64p]
    lib.mocb200_set_source_xs.argtypes = [H, C.c_int, i32p, _f64p, _f64p, _f64p, i32p]
    lib.mocb200_set_external_source.argtypes = [H, _f64p]
    lib.mocb200_fission_source.argtypes = [H, C.c_double]
    lib.mocb200_set_fission_source.argtypes = [H, _f64p]
    lib.mocb200_get_fission_source.argtypes = [H, _f64p]
    lib.mocb200_build_source.argtypes = [H, C.c_int, C.c_int]
    lib.mocb200_angle_families.argtypes = [C.POINTER(Problem), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.mocb200_sweep_partial.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.mocb200_finalize_flux.argtypes = [H, C.c_int, C.c_int]
    lib.mocb200_device_buffer.argtypes = [H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    lib.mocb200_adopt_device_buffer.argtypes = [H, C.c_int, C.c_void_p, C.c_int64]
    lib.mocb200_version.restype = C.c_char_p
    if path == LIB_PATH:
        _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(_f64p)


BUF_TALLY, BUF_CURRENT, BUF_SURFACE_FLUX = 0, 1, 2


def material_tables(xs_nf, xs_ch, xs_scat):
    """Cross-section-mesh regions out of per-FSR arrays (what a flattened problem carries): xs_nf / xs_ch [G][n_reg],
    xs_scat [G to][G from][n_reg] -> (fsr_mat [n_reg] int32, xsnf [n_mat][G], xsch [n_mat][G], scat [n_mat][to][from]).
    The C++ plugin passes the reference's XSMesh regions directly."""
    xs_nf, xs_ch, xs_scat = (np.asarray(a, dtype=np.float64) for a in (xs_nf, xs_ch, xs_scat))
    G, n_reg = xs_nf.shape
    cols = np.concatenate([xs_nf, xs_ch, xs_scat.reshape(G * G, n_reg)], axis=0).T  # one row per FSR
    uniq, inv = np.unique(cols, axis=0, return_inverse=True)
    return (np.ascontiguousarray(inv.reshape(-1), dtype=np.int32), np.ascontiguousarray(uniq[:, :G]),
            np.ascontiguousarray(uniq[:, G:2 * G]), np.ascontiguousarray(uniq[:, 2 * G:].reshape(-1, G, G)))


def angle_families(arrays, lib=None):
    """(number of angle families, family of every boundary angle [2 n_ang]) of a flattened problem -- the units
    a single plane is sharded by (mocb200_angle_families; needs no device)."""
    lib = lib or load_library()
    prob, keep = problem_from_arrays(arrays)
    n = C.c_int32()
    fam = np.zeros(2 * int(prob.n_ang), dtype=np.int32)
    rc = lib.mocb200_angle_families(C.byref(prob), C.byref(n), fam.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise RuntimeError(f"mocb200_angle_families failed ({rc})")
    return n.value, fam


def _host(a, shape):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
    return a


class Sweeper:
    """Handle on a device-resident MoC problem (one GPU)."""

    def __init__(self, arrays, device=0, boundary_update=BOUNDARY_GS, exp_mode=EXP_TABLE, max_polar=0,
                 block_threads=0, plane_begin=0, plane_end=0, kernel=0, chunk_cap=0, cache_groups=0, persistent=0,
                 family_begin=0, family_end=0, lib=None):
        self.lib = lib or load_library()
        self.arrays = arrays
        self.problem, self._keep = problem_from_arrays(arrays)
        self.n_group = self.problem.n_group
        self.n_reg = self.problem.n_reg
        self.n_plane = self.problem.n_plane
        self.n_surf = self.problem.n_surf
        self.bc_per_group = self.problem.bc_per_group
        opt = Options()
        opt.device, opt.boundary_update, opt.exp_mode = device, boundary_update, exp_mode
        opt.max_polar, opt.block_threads = max_polar, block_threads
        opt.plane_begin, opt.plane_end = plane_begin, plane_end
        opt.kernel = kernel
        opt.chunk_cap = chunk_cap
        opt.cache_groups = cache_groups
        opt.persistent = persistent
        opt.family_begin, opt.family_end = family_begin, family_end
        self.h = C.c_void_p()
        rc = self.lib.mocb200_create(C.byref(self.problem), C.byref(opt), C.byref(self.h))
        if rc != 0:
            msg = self.lib.mocb200_last_error(None).decode()
            self.h = None
            raise RuntimeError(f"mocb200_create failed ({rc}): {msg}")

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.mocb200_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.mocb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        self._ck(self.lib.mocb200_set_stream(self.h, C.c_void_p(cuda_stream)), "set_stream")

    def synchronize(self):
        self._ck(self.lib.mocb200_synchronize(self.h), "synchronize")

    def set_xs(self, g_begin, xstr, xstr_src=None, xs_self=None):
        xstr = np.ascontiguousarray(xstr, dtype=np.float64).reshape(-1, self.n_reg)
        gc = xstr.shape[0]
        src = None if xstr_src is None else _host(np.reshape(xstr_src, (-1, self.n_reg)), (gc, self.n_reg))
        if xs_self is None:
            xs_self = np.zeros((gc, self.n_reg))
        slf = _host(np.reshape(xs_self, (-1, self.n_reg)), (gc, self.n_reg))
        self._ck(self.lib.mocb200_set_xs(self.h, g_begin, gc, _ptr(xstr),
                                         _ptr(src) if src is not None else None, _ptr(slf)), "set_xs")

    def _set(self, fn, g_begin, a):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, self.n_reg)
        self._ck(getattr(self.lib, fn)(self.h, g_begin, a.shape[0], _ptr(a)), fn)

    def set_source(self, g_begin, src):
        self._set("mocb200_set_source", g_begin, src)

    def set_flux(self, g_begin, flux):
        self._set("mocb200_set_flux", g_begin, flux)

    def set_qbar(self, g_begin, qbar):
        self._set("mocb200_set_qbar", g_begin, qbar)

    def get_flux(self, g_begin, g_count, out=None):
        if out is None:
            out = np.empty((g_count, self.n_reg))
        self._ck(self.lib.mocb200_get_flux(self.h, g_begin, g_count, _ptr(out)), "get_flux")
        return out

    def set_boundary(self, plane, g_begin, bc):
        bc = np.ascontiguousarray(bc, dtype=np.float64).reshape(-1, self.bc_per_group)
        self._ck(self.lib.mocb200_set_boundary(self.h, plane, g_begin, bc.shape[0], _ptr(bc)), "set_boundary")

    def get_boundary(self, plane, g_begin, g_count):
        out = np.empty((g_count, self.bc_per_group))
        self._ck(self.lib.mocb200_get_boundary(self.h, plane, g_begin, g_count, _ptr(out)), "get_boundary")
        return out

    def sweep(self, g_begin, g_count, n_inner=1, tally_mode=TALLY_NONE, use_qbar=False):
        self._ck(self.lib.mocb200_sweep(self.h, g_begin, g_count, n_inner, tally_mode, int(use_qbar)), "sweep")

    # ---- source construction on the device (FixedSourceSolver::step / calc_fission_source) ----
    def set_source_xs(self, fsr_mat, xsnf, xsch, scat, band=None):
        fsr_mat = np.ascontiguousarray(fsr_mat, dtype=np.int32)
        xsnf, xsch, scat = (np.ascontiguousarray(a, dtype=np.float64) for a in (xsnf, xsch, scat))
        n_mat, G = xsnf.shape
        assert fsr_mat.size == self.n_reg and G == self.n_group and xsch.shape == (n_mat, G) and scat.shape == (n_mat, G, G)
        i32p = C.POINTER(C.c_int32)
        if band is not None:
            band = np.ascontiguousarray(band, dtype=np.int32)
            assert band.shape == (n_mat, G, 2)
        self._ck(self.lib.mocb200_set_source_xs(self.h, n_mat, fsr_mat.ctypes.data_as(i32p), _ptr(xsnf), _ptr(xsch),
                                                _ptr(scat), band.ctypes.data_as(i32p) if band is not None else None),
                 "set_source_xs")

    def set_external_source(self, ext):
        if ext is None:
            self._ck(self.lib.mocb200_set_external_source(self.h, None), "set_external_source")
        else:
            ext = _host(ext, (self.n_group, self.n_reg))
            self._ck(self.lib.mocb200_set_external_source(self.h, _ptr(ext)), "set_external_source")

    def fission_source(self, k):
        self._ck(self.lib.mocb200_fission_source(self.h, float(k)), "fission_source")

    def set_fission_source(self, fs):
        fs = _host(fs, (self.n_reg,))
        self._ck(self.lib.mocb200_set_fission_source(self.h, _ptr(fs)), "set_fission_source")

    def get_fission_source(self):
        fs = np.empty(self.n_reg)
        self._ck(self.lib.mocb200_get_fission_source(self.h, _ptr(fs)), "get_fission_source")
        return fs

    def build_source(self, g_begin, g_count=1):
        self._ck(self.lib.mocb200_build_source(self.h, g_begin, g_count), "build_source")

    def get_source(self, g_begin, g_count):
        out = np.empty((g_count, self.n_reg))
        self._ck(self.lib.mocb200_get_source(self.h, g_begin, g_count, _ptr(out)), "get_source")
        return out

    def sweep_partial(self, g_begin, g_count=1, tally_mode=TALLY_NONE, use_qbar=False):
        """One inner sweep of the handle's angle families; the tally stays un-normalised (sum it over the ranks,
        then finalize_flux)."""
        self._ck(self.lib.mocb200_sweep_partial(self.h, g_begin, g_count, tally_mode, int(use_qbar)), "sweep_partial")

    def finalize_flux(self, g_begin, g_count=1):
        self._ck(self.lib.mocb200_finalize_flux(self.h, g_begin, g_count), "finalize_flux")

    def device_buffer(self, which):
        """(device address, doubles) of BUF_TALLY / BUF_CURRENT / BUF_SURFACE_FLUX."""
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self.lib.mocb200_device_buffer(self.h, which, C.byref(p), C.byref(n)), "device_buffer")
        return p.value, n.value

    def adopt_device_buffer(self, which, device_ptr, count):
        self._ck(self.lib.mocb200_adopt_device_buffer(self.h, which, C.c_void_p(device_ptr), count), "adopt_device_buffer")

    def get_coarse(self, group):
        cur = np.zeros(self.n_surf)
        sf = np.zeros(self.n_surf)
        self._ck(self.lib.mocb200_get_coarse(self.h, group, _ptr(cur), _ptr(sf)), "get_coarse")
        return cur, sf

    def set_sn_xs(self, g_begin, xs):
        n = self.n_plane * int(self.problem.n_cell_plane)
        xs = np.ascontiguousarray(xs, dtype=np.float64).reshape(-1, n)
        self._ck(self.lib.mocb200_set_sn_xs(self.h, g_begin, xs.shape[0], _ptr(xs)), "set_sn_xs")

    def get_corrections(self, group, alpha=None, beta=None):
        """alpha [2*n_ang, n_cell_total, 2], beta [2*n_ang, n_cell_total] of the last corrections sweep."""
        n_ang2 = 2 * int(self.problem.n_ang)
        n = self.n_plane * int(self.problem.n_cell_plane)
        if alpha is None:
            alpha = np.full((n_ang2, n, 2), np.nan)
        if beta is None:
            beta = np.full((n_ang2, n), np.nan)
        self._ck(self.lib.mocb200_get_corrections(self.h, group, _ptr(alpha), _ptr(beta)), "get_corrections")
        return alpha, beta

    def set_sweep_inputs(self, group, source=None, flux=None, boundary=None):
        """Fused upload for one sweep(group): source[n_reg], flux[n_reg], boundary = list of per-plane
        [bc_per_group] arrays (None entries skipped). One staged copy, no synchronisation."""
        keep = []

        def arr(a, n):
            if a is None:
                return None
            a = _host(a, (n,))
            keep.append(a)
            return _ptr(a)
        bptr = None
        if boundary is not None:
            bptr = (_f64p * self.n_plane)()
            for ip, b in enumerate(boundary):
                if b is not None:
                    bptr[ip] = arr(b, self.bc_per_group)
        self._ck(self.lib.mocb200_set_sweep_inputs(self.h, group, arr(source, self.n_reg), arr(flux, self.n_reg), bptr),
                 "set_sweep_inputs")

    def get_sweep_results(self, group, flux=None, boundary=None, coarse=False):
        """Fused download after one sweep(group): fills flux[n_reg] and the per-plane boundary arrays in place
        (None = skip) and returns (current, surface_flux) when coarse; one synchronisation."""
        bptr = None
        if boundary is not None:
            bptr = (_f64p * self.n_plane)()
            for ip, b in enumerate(boundary):
                if b is not None:
                    assert b.dtype == np.float64 and b.flags.c_contiguous and b.size == self.bc_per_group
                    bptr[ip] = _ptr(b)
        if flux is not None:
            assert flux.dtype == np.float64 and flux.flags.c_contiguous and flux.size == self.n_reg
        cur = np.zeros(self.n_surf) if coarse else None  # a handle fills the surfaces of its own planes
        sf = np.zeros(self.n_surf) if coarse else None
        self._ck(self.lib.mocb200_get_sweep_results(self.h, group, _ptr(flux) if flux is not None else None, bptr,
                                                    _ptr(cur) if coarse else None, _ptr(sf) if coarse else None),
                 "get_sweep_results")
        return cur, sf

    def pack_results_device(self, group, device_ptr=None, capacity=0):
        """Packs [flux of the handle's FSR range | coarse current | surface flux of its planes] into a device buffer
        (raw pointer, e.g. tensor.data_ptr()) on the handle's stream; returns the doubles written (size query with
        device_ptr None)."""
        n = C.c_size_t()
        self._ck(self.lib.mocb200_pack_results_device(self.h, group, C.c_void_p(device_ptr) if device_ptr else None,
                                                      C.c_size_t(capacity), C.byref(n)), "pack_results_device")
        return n.value

    def stats(self):
        s = Stats()
        self._ck(self.lib.mocb200_get_stats(self.h, C.byref(s)), "get_stats")
        return {"kernel_launches": s.kernel_launches, "sweep_launches": s.sweep_launches,
                "segments_per_sweep": s.segments_per_sweep, "unique_segments": s.unique_segments,
                "device_bytes": s.device_bytes, "items": [s.items[0], s.items[1]], "kernel": s.kernel,
                "swept_segments": s.swept_segments}

    def set_timing(self, enabled=True):
        self._ck(self.lib.mocb200_set_timing(self.h, int(enabled)), "set_timing")

    def get_timing(self):
        """(summed sweep-kernel ms, inner sweeps) since the last call."""
        ms, n = C.c_double(), C.c_int64()
        self._ck(self.lib.mocb200_get_timing(self.h, C.byref(ms), C.byref(n)), "get_timing")
        return ms.value, n.value

    def last_sweep_ms(self):
        ms = C.c_double()
        self._ck(self.lib.mocb200_last_sweep_ms(self.h, C.byref(ms)), "last_sweep_ms")
        return ms.value
