"""GPU parity tests: the CUDA sweep (through the C ABI) against the reference's golden
records and against the CPU oracle on the same inputs.

Tolerances: boundary (angular) flux, scalar flux and coarse tallies are FP64 and differ
from the reference only by FMA contraction and by the order in which the atomics land:
relative 1e-11 (north_star asks 1e-5 on flux). Index work (FSR ids, boundary linkage,
coarse-surface linkage) is exercised implicitly: any mismatch is an O(1) error.
"""
import numpy as np
import pytest

from conftest import CASES, load_case, records
from oracle_lib import oracle_sweep1g

pytestmark = pytest.mark.gpu

RTOL = 1e-11


def _sweeper(flat, **kw):
    from mocc_b200 import Sweeper
    return Sweeper(flat, **kw)


def _close(a, b, rtol=RTOL, atol=None):
    atol = 1e-14 if atol is None else atol
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), atol / rtol)
    err = np.max(np.abs(a - b) / scale) if a.size else 0.0
    assert err < rtol, f"max rel err {err:.3e}"


def _run_record(sw, flat, rec, use_qbar=True, gold=None):
    g = int(rec["group"][0])
    mode = int(rec["mode"][0])
    n_plane = sw.n_plane
    if use_qbar:
        sw.set_xs(g, rec["xstr"])
        sw.set_qbar(g, rec["qbar"])
    else:
        sw.set_xs(g, rec["xstr"], xstr_src=gold[f"xs_tr_{g}"], xs_self=gold[f"xs_self_{g}"])
        sw.set_source(g, rec["src"])
        sw.set_flux(g, rec["flux_in"])
    bc = rec["bc_in"].reshape(n_plane, sw.bc_per_group)
    for ip in range(n_plane):
        sw.set_boundary(ip, g, bc[ip])
    sw.sweep(g, 1, n_inner=1, tally_mode=mode, use_qbar=use_qbar)
    flux = sw.get_flux(g, 1)[0]
    bc_out = np.concatenate([sw.get_boundary(ip, g, 1)[0] for ip in range(n_plane)])
    cur = sf = None
    if mode == 1:
        cur, sf = sw.get_coarse(g)
    return flux, bc_out, cur, sf


def _xy_mask(flat):
    n_surf, nsp = int(flat["n_surf"][0]), int(flat["n_surf_plane"][0])
    nxy = int(flat["nx"][0]) * int(flat["ny"][0])
    m = np.zeros(n_surf, dtype=bool)
    for off in flat["plane_surf_offset"]:
        m[off + nxy: off + nsp] = True
    return m


@pytest.mark.parametrize("case", [c for c in sorted(CASES) if c != "mini3d_2d3d"])
@pytest.mark.parametrize("max_polar", [1, 2, 4])
@pytest.mark.parametrize("kernel", [1, 2, 3, 4, 5])
def test_sweep1g_matches_reference_golden(case, max_polar, kernel):
    flat, gold = load_case(case)
    gs = bool(gold["gs_boundary"][0])
    sw = _sweeper(flat, boundary_update=0 if gs else 1, max_polar=max_polar, kernel=kernel)
    xy = _xy_mask(flat)
    for rec in records(gold):
        flux, bc_out, cur, sf = _run_record(sw, flat, rec)
        _close(flux, rec["flux_out"])
        _close(bc_out, rec["bc_out"])
        if cur is not None:
            area = flat["surf_area"]
            _close(cur[xy] / area[xy], rec["current"][xy], atol=1e-13)
            _close(sf[xy] / area[xy], rec["surface_flux"][xy], atol=1e-13)
    sw.close()


@pytest.mark.parametrize("case", ["mini2d_gs", "mini2d_jacobi", "mini3d_gs", "3x3_s05_gs"])
@pytest.mark.parametrize("max_polar", [1, 2, 3])
@pytest.mark.parametrize("chunk_cap", [32, -64, -100000])
def test_chunk_kernel_superblocks_match_reference_golden(case, max_polar, chunk_cap):
    """CHUNK kernel with a small staging cap: every longer track is chained through super-blocks
    (negative cap: two warps per track, with and without super-blocks)."""
    flat, gold = load_case(case)
    gs = bool(gold["gs_boundary"][0])
    sw = _sweeper(flat, boundary_update=0 if gs else 1, max_polar=max_polar, kernel=4, chunk_cap=chunk_cap)
    xy = _xy_mask(flat)
    for rec in records(gold):
        flux, bc_out, cur, sf = _run_record(sw, flat, rec)
        _close(flux, rec["flux_out"])
        _close(bc_out, rec["bc_out"])
        if cur is not None:
            area = flat["surf_area"]
            _close(cur[xy] / area[xy], rec["current"][xy], atol=1e-13)
            _close(sf[xy] / area[xy], rec["surface_flux"][xy], atol=1e-13)
    sw.close()


@pytest.mark.parametrize("case", ["mini2d_gs", "mini2d_jacobi", "mini3d_gs", "3x3_s05_gs"])
@pytest.mark.parametrize("max_polar", [1, 2, 4])
@pytest.mark.parametrize("chunk_cap", [14, 40, 100])
def test_rchunk_kernel_chained_units_match_reference_golden(case, max_polar, chunk_cap):
    """RCHUNK kernel with few chunks per batch: every longer track becomes a chained unit (sub-blocks linked by a
    carried flux, two passes), single-batch and chained units interleaved in the work list."""
    flat, gold = load_case(case)
    gs = bool(gold["gs_boundary"][0])
    sw = _sweeper(flat, boundary_update=0 if gs else 1, max_polar=max_polar, kernel=5, chunk_cap=chunk_cap)
    xy = _xy_mask(flat)
    for rec in records(gold):
        flux, bc_out, cur, sf = _run_record(sw, flat, rec)
        _close(flux, rec["flux_out"])
        _close(bc_out, rec["bc_out"])
        if cur is not None:
            area = flat["surf_area"]
            _close(cur[xy] / area[xy], rec["current"][xy], atol=1e-13)
            _close(sf[xy] / area[xy], rec["surface_flux"][xy], atol=1e-13)
    sw.close()


@pytest.mark.parametrize("case", ["mini2d_gs", "mini3d_gs", "3x3_s05_gs"])
def test_self_scatter_then_sweep_matches_reference(case):
    flat, gold = load_case(case)
    sw = _sweeper(flat, boundary_update=0)
    for rec in records(gold):
        flux, bc_out, _, _ = _run_record(sw, flat, rec, use_qbar=False, gold=gold)
        _close(flux, rec["flux_out"])
        _close(bc_out, rec["bc_out"])
    sw.close()


@pytest.mark.parametrize("case", ["mini2d_gs", "mini2d_jacobi", "mini3d_gs"])
def test_fused_transfers_equal_the_single_calls(case):
    """mocb200_set_sweep_inputs / get_sweep_results == set_source + set_flux + set_boundary / get_flux +
    get_boundary + get_coarse (what the plugin moves around every sweep(group))."""
    flat, gold = load_case(case)
    gs = bool(gold["gs_boundary"][0])
    n_reg, bcpg = int(flat["n_reg"][0]), int(flat["bc_per_group"][0])
    res = []
    for fused in (False, True):
        sw = _sweeper(flat, boundary_update=0 if gs else 1)
        n_plane = sw.n_plane
        out = []
        for rec in records(gold):
            g = int(rec["group"][0])
            sw.set_xs(g, rec["xstr"], xstr_src=gold[f"xs_tr_{g}"], xs_self=gold[f"xs_self_{g}"])
            bc = rec["bc_in"].reshape(n_plane, bcpg)
            if fused:
                sw.set_sweep_inputs(g, rec["src"], rec["flux_in"], [bc[ip] for ip in range(n_plane)])
            else:
                sw.set_source(g, rec["src"])
                sw.set_flux(g, rec["flux_in"])
                for ip in range(n_plane):
                    sw.set_boundary(ip, g, bc[ip])
            sw.sweep(g, 1, n_inner=2, tally_mode=1)
            if fused:
                flux = np.zeros(n_reg)
                bo = [np.zeros(bcpg) for _ in range(n_plane)]
                cur, sf = sw.get_sweep_results(g, flux, bo, coarse=True)
                bo = np.concatenate(bo)
            else:
                flux = sw.get_flux(g, 1)[0]
                bo = np.concatenate([sw.get_boundary(ip, g, 1)[0] for ip in range(n_plane)])
                cur, sf = sw.get_coarse(g)
            out.append((flux, bo, cur, sf))
        sw.close()
        res.append(out)
    for a, b in zip(*res):
        for x, y in zip(a, b):
            _close(x, y)


@pytest.mark.parametrize("case", ["mini2d_gs", "mini3d_gs"])
@pytest.mark.parametrize("jacobi", [False, True])
@pytest.mark.parametrize("kernel", [1, 2, 3])
def test_batched_groups_match_oracle(case, jacobi, kernel):
    """All groups in one launch (groups across lanes) == the oracle run group by group."""
    flat, gold = load_case(case)
    G, n_reg, n_plane = (int(flat[k][0]) for k in ("n_group", "n_reg", "n_plane"))
    bcpg = int(flat["bc_per_group"][0])
    rng = np.random.default_rng(7)
    xstr = np.stack([gold[f"xs_tr_{g}"] for g in range(G)])
    qbar = rng.uniform(0.05, 1.0, size=(G, n_reg))
    bc = rng.uniform(0.0, 0.3, size=(n_plane, G, bcpg))
    sw = _sweeper(flat, boundary_update=1 if jacobi else 0, kernel=kernel)
    sw.set_xs(0, xstr)
    sw.set_qbar(0, qbar)
    for ip in range(n_plane):
        sw.set_boundary(ip, 0, bc[ip])
    sw.sweep(0, G, n_inner=1, tally_mode=1, use_qbar=True)
    flux = sw.get_flux(0, G)
    xy = _xy_mask(flat)
    for g in range(G):
        f_o, bc_o, cur_o, sf_o = oracle_sweep1g(flat, xstr[g], qbar[g], bc[:, g, :], gs_boundary=not jacobi,
                                                tally_mode=1)
        _close(flux[g], f_o)
        bc_g = np.stack([sw.get_boundary(ip, g, 1)[0] for ip in range(n_plane)])
        _close(bc_g, bc_o)
        cur, sf = sw.get_coarse(g)
        area = flat["surf_area"]
        _close(cur[xy] / area[xy], cur_o[xy], atol=1e-13)
        _close(sf[xy] / area[xy], sf_o[xy], atol=1e-13)
    sw.close()


@pytest.mark.parametrize("kernel", [1, 2, 3, 4, 5])
def test_exponential_arguments_outside_the_table_match_oracle(kernel):
    """Exponential_Linear falls back to std::exp outside [-10, 0] (exponential.hpp:71-75): an optically thick
    region (tau > 10) and a NEGATIVE transport cross section (possible under transverse-leakage splitting:
    argument > 0) must both go through that branch, not through an out-of-range table read."""
    flat, gold = load_case("mini2d_gs")
    n_reg, n_plane, bcpg = int(flat["n_reg"][0]), int(flat["n_plane"][0]), int(flat["bc_per_group"][0])
    rng = np.random.default_rng(11)
    xstr = gold["xs_tr_0"].copy()
    xstr[::7] = 60.0    # tau up to ~70 per segment: below the table
    xstr[3::11] = -0.02  # above the table
    qbar = rng.uniform(0.05, 1.0, size=n_reg)
    bc = rng.uniform(0.0, 0.3, size=(n_plane, bcpg))
    sw = _sweeper(flat, boundary_update=0, kernel=kernel)
    sw.set_xs(0, xstr)
    sw.set_qbar(0, qbar)
    for ip in range(n_plane):
        sw.set_boundary(ip, 0, bc[ip])
    sw.sweep(0, 1, n_inner=1, tally_mode=0, use_qbar=True)
    f_o, bc_o, _, _ = oracle_sweep1g(flat, xstr, qbar, bc, gs_boundary=True, tally_mode=0)
    _close(sw.get_flux(0, 1)[0], f_o, rtol=1e-10)
    _close(np.stack([sw.get_boundary(ip, 0, 1)[0] for ip in range(n_plane)]), bc_o, rtol=1e-10)
    sw.close()


@pytest.mark.parametrize("kernel", [2, 3, 4, -4, 5, -5])
@pytest.mark.parametrize("max_polar", [1, 2])
def test_corrections_match_reference_golden(kernel, max_polar):
    """MOCB200_TALLY_CORRECTIONS == the reference's MoCSweeper_2D3D last inner (cmdo::CurrentCorrections):
    flux, boundary flux, coarse currents / surface flux (backward surface flux SUBTRACTED, the worker's quirk)
    and the alpha/beta correction factors."""
    flat, gold = load_case("mini3d_2d3d")
    recs = [r for r in records(gold) if int(r["mode"][0]) == 2]
    assert recs
    # kernel -4: the chunk kernel with a 64-segment staging cap and two-warp teams (super-block chaining)
    # kernel -5: the register-chunk kernel with three chunks per batch (chained units)
    kw = dict(kernel=4, chunk_cap=-64) if kernel == -4 else (dict(kernel=5, chunk_cap=40) if kernel == -5 else dict(kernel=kernel))
    sw = _sweeper(flat, boundary_update=0, max_polar=max_polar, **kw)
    xy = _xy_mask(flat)
    area = flat["surf_area"]
    n_plane = sw.n_plane
    for rec in recs:
        g = int(rec["group"][0])
        sw.set_xs(g, rec["xstr"], xstr_src=rec["xstr_true"], xs_self=gold[f"xs_self_{g}"])
        sw.set_qbar(g, rec["qbar"])
        sw.set_sn_xs(g, rec["sn_xs"])
        bc = rec["bc_in"].reshape(n_plane, sw.bc_per_group)
        for ip in range(n_plane):
            sw.set_boundary(ip, g, bc[ip])
        sw.sweep(g, 1, n_inner=1, tally_mode=2, use_qbar=True)
        _close(sw.get_flux(g, 1)[0], rec["flux_out"])
        bc_out = np.concatenate([sw.get_boundary(ip, g, 1)[0] for ip in range(n_plane)])
        _close(bc_out, rec["bc_out"])
        cur, sf = sw.get_coarse(g)
        _close(cur[xy] / area[xy], rec["current"][xy], atol=1e-13)
        _close(sf[xy] / area[xy], rec["surface_flux"][xy], atol=1e-13)
        alpha, beta = sw.get_corrections(g)
        _close(alpha.ravel(), rec["alpha"], rtol=1e-10)
        _close(beta.ravel(), rec["beta"], rtol=1e-10)
    sw.close()


def test_corrections_batched_groups_match_oracle():
    """Correction factors with all groups in one batch (8 group lanes) == the oracle group by group."""
    from oracle_lib import oracle_sweep1g_corrections
    flat, gold = load_case("mini3d_2d3d")
    G, n_reg, n_plane = (int(flat[k][0]) for k in ("n_group", "n_reg", "n_plane"))
    bcpg, ncp = int(flat["bc_per_group"][0]), int(flat["n_cell_plane"][0])
    rng = np.random.default_rng(11)
    xstr = np.stack([gold[f"xs_tr_{g}"] for g in range(G)])
    xsplit = xstr * rng.uniform(1.0, 1.2, size=xstr.shape)
    qbar = rng.uniform(0.05, 1.0, size=(G, n_reg))
    sn = rng.uniform(0.3, 1.5, size=(G, n_plane * ncp))
    bc = rng.uniform(0.0, 0.3, size=(n_plane, G, bcpg))
    sw = _sweeper(flat, boundary_update=0)
    sw.set_xs(0, xsplit, xstr_src=xstr, xs_self=np.zeros_like(xstr))
    sw.set_qbar(0, qbar)
    sw.set_sn_xs(0, sn)
    for ip in range(n_plane):
        sw.set_boundary(ip, 0, bc[ip])
    sw.sweep(0, G, n_inner=1, tally_mode=2, use_qbar=True)
    flux = sw.get_flux(0, G)
    for g in range(G):
        f_o, bc_o, cur_o, sf_o, al_o, be_o = oracle_sweep1g_corrections(flat, xsplit[g], xstr[g], qbar[g], sn[g],
                                                                        bc[:, g, :], gs_boundary=True)
        _close(flux[g], f_o)
        alpha, beta = sw.get_corrections(g)
        _close(alpha, al_o, rtol=1e-10)
        _close(beta, be_o, rtol=1e-10)
    sw.close()


def test_inner_iterations_match_oracle():
    """n_inner inner iterations on the device == oracle self-scatter + sweep repeated."""
    from oracle_lib import oracle_self_scatter
    flat, gold = load_case("mini2d_gs")
    rec = records(gold)[0]
    g = int(rec["group"][0])
    sw = _sweeper(flat, boundary_update=0)
    sw.set_xs(g, rec["xstr"], xstr_src=gold[f"xs_tr_{g}"], xs_self=gold[f"xs_self_{g}"])
    sw.set_source(g, rec["src"])
    sw.set_flux(g, rec["flux_in"])
    sw.set_boundary(0, g, rec["bc_in"])
    sw.sweep(g, 1, n_inner=3)
    flux = sw.get_flux(g, 1)[0]
    f, bc = rec["flux_in"], rec["bc_in"].reshape(1, -1)
    for _ in range(3):
        q = oracle_self_scatter(rec["src"], f, gold[f"xs_self_{g}"], gold[f"xs_tr_{g}"])
        f, bc, _, _ = oracle_sweep1g(flat, rec["xstr"], q, bc, gs_boundary=True)
    _close(flux, f)
    _close(sw.get_boundary(0, g, 1)[0], bc[0])
    sw.close()


def test_plane_sharding_equals_whole():
    """Two handles owning disjoint macroplane ranges reproduce the single-handle result."""
    flat, gold = load_case("mini3d_gs")
    rec = records(gold)[0]
    g = int(rec["group"][0])
    n_plane = int(flat["n_plane"][0])
    whole = _sweeper(flat)
    f_w, bc_w, _, _ = _run_record(whole, flat, rec)
    parts = [_sweeper(flat, plane_begin=0, plane_end=1), _sweeper(flat, plane_begin=1, plane_end=n_plane)]
    first = list(flat["plane_first_reg"]) + [int(flat["n_reg"][0])]
    flux = np.zeros_like(f_w)
    bcpg = parts[0].bc_per_group
    for sw, (lo, hi) in zip(parts, [(0, 1), (1, n_plane)]):
        f, bc, _, _ = _run_record(sw, flat, rec)
        flux[first[lo]:first[hi]] = f[first[lo]:first[hi]]
        assert np.array_equal(bc[lo * bcpg:hi * bcpg], bc_w[lo * bcpg:hi * bcpg]) or \
            np.allclose(bc[lo * bcpg:hi * bcpg], bc_w[lo * bcpg:hi * bcpg], rtol=1e-13)
    _close(flux, f_w)


def test_errors_are_reported():
    flat, _ = load_case("mini2d_gs")
    sw = _sweeper(flat)
    with pytest.raises(RuntimeError, match="set_xs"):
        sw.sweep(0, 1)
    with pytest.raises(RuntimeError, match="group range"):
        sw.set_qbar(99, np.zeros(sw.n_reg))
    sw.close()


@pytest.fixture(scope="module")
def c5g7_2d_flat():
    """BASELINE.json config 2 at full size, flattened on the box by the plugin's own set-up code."""
    import bench
    from mocc_b200 import load_arrays
    return load_arrays(bench.workload_files())


@pytest.mark.parametrize("kernel,jacobi", [(0, False), (0, True), (3, False), (4, False)])
def test_full_size_c5g7_2d_sweeps_match_oracle(c5g7_2d_flat, kernel, jacobi):
    """Full-size parity (18.2 M reference segments per group sweep): every group, with and without the coarse-current
    tally, two inners with self scatter on the device, against the C oracle on the same seeded inputs."""
    flat = c5g7_2d_flat
    G, n_reg, bcpg = (int(flat[k][0]) for k in ("n_group", "n_reg", "bc_per_group"))
    rng = np.random.default_rng(2025)
    sw = _sweeper(flat, boundary_update=1 if jacobi else 0, kernel=kernel)
    assert sw.stats()["kernel"] == (5 if kernel == 0 else kernel)
    sw.set_xs(0, flat["xs_tr"], xstr_src=flat["xs_tr"], xs_self=flat["xs_self"])
    xy = _xy_mask(flat)
    area = flat["surf_area"]
    for g in range(G):
        tally = g % 2
        qbar = rng.uniform(0.05, 1.0, size=n_reg)
        bc = rng.uniform(0.0, 0.3, size=(1, bcpg))
        sw.set_qbar(g, qbar)
        sw.set_boundary(0, g, bc[0])
        sw.sweep(g, 1, n_inner=1, tally_mode=tally, use_qbar=True)
        f_o, bc_o, cur_o, sf_o = oracle_sweep1g(flat, flat["xs_tr"][g], qbar, bc, gs_boundary=not jacobi, tally_mode=tally)
        _close(sw.get_flux(g, 1)[0], f_o)
        _close(sw.get_boundary(0, g, 1)[0], bc_o[0])
        if tally:
            cur, sf = sw.get_coarse(g)
            _close(cur[xy] / area[xy], cur_o[xy], atol=1e-13)
            _close(sf[xy] / area[xy], sf_o[xy], atol=1e-13)
    sw.close()


def test_full_size_sweep_is_linear_in_source_and_boundary_flux(c5g7_2d_flat):
    """Size-independent property: for fixed cross sections the sweep is linear in (q-bar, incoming flux)."""
    flat = c5g7_2d_flat
    n_reg, bcpg = int(flat["n_reg"][0]), int(flat["bc_per_group"][0])
    rng = np.random.default_rng(7)
    g = 5
    sw = _sweeper(flat, boundary_update=0)
    sw.set_xs(0, flat["xs_tr"], xstr_src=flat["xs_tr"], xs_self=flat["xs_self"])
    q = [rng.uniform(0.05, 1.0, size=n_reg) for _ in range(2)]
    b = [rng.uniform(0.0, 0.3, size=bcpg) for _ in range(2)]
    out = []
    for qq, bb in ((q[0], b[0]), (q[1], b[1]), (2.0 * q[0] + 0.5 * q[1], 2.0 * b[0] + 0.5 * b[1])):
        sw.set_qbar(g, qq)
        sw.set_boundary(0, g, bb)
        sw.sweep(g, 1, n_inner=1, tally_mode=1, use_qbar=True)
        out.append((sw.get_flux(g, 1)[0], sw.get_boundary(0, g, 1)[0]) + sw.get_coarse(g))
    for x0, x1, x2 in zip(*out):
        _close(x2, 2.0 * x0 + 0.5 * x1, rtol=1e-10, atol=1e-12)
    sw.close()


@pytest.mark.parametrize("case", ["mini2d_gs", "mini3d_gs"])
@pytest.mark.parametrize("n_inner", [1, 3])
@pytest.mark.parametrize("kernel", [4, 5])
def test_two_groups_per_call_on_the_chunk_kernel(case, n_inner, kernel):
    """g_count = 2 keeps one group per warp (chunk kernel, group-major q-bar / tally): equals two single-group calls."""
    flat, gold = load_case(case)
    G, n_reg, n_plane = (int(flat[k][0]) for k in ("n_group", "n_reg", "n_plane"))
    bcpg = int(flat["bc_per_group"][0])
    rng = np.random.default_rng(5)
    xstr = np.stack([gold[f"xs_tr_{g}"] for g in range(G)])
    xself = np.stack([gold[f"xs_self_{g}"] for g in range(G)])
    src = rng.uniform(0.05, 1.0, size=(G, n_reg))
    flux0 = rng.uniform(0.5, 1.5, size=(G, n_reg))
    bc = rng.uniform(0.0, 0.3, size=(n_plane, G, bcpg))
    res = []
    for pair in (True, False):
        sw = _sweeper(flat, boundary_update=0, kernel=kernel)
        sw.set_xs(0, xstr, xstr_src=xstr, xs_self=xself)
        sw.set_source(0, src)
        sw.set_flux(0, flux0)
        for ip in range(n_plane):
            sw.set_boundary(ip, 0, bc[ip])
        if pair:
            sw.sweep(0, 2, n_inner=n_inner, tally_mode=1)
        else:
            for g in range(2):
                sw.sweep(g, 1, n_inner=n_inner, tally_mode=1)
        out = [sw.get_flux(0, 2)]
        for g in range(2):
            out.append(np.stack([sw.get_boundary(ip, g, 1)[0] for ip in range(n_plane)]))
            out.extend(sw.get_coarse(g))
        res.append(out)
        sw.close()
    for x, y in zip(*res):
        _close(x, y, atol=1e-13)


@pytest.mark.parametrize("persistent", [1, 2])
@pytest.mark.parametrize("jacobi", [False, True])
@pytest.mark.parametrize("n_inner,tally", [(1, 0), (3, 0), (4, 1)])
@pytest.mark.parametrize("case", ["mini2d_gs", "3x3_s05_gs"])
def test_persistent_launch_equals_per_phase_launches(case, persistent, jacobi, n_inner, tally):
    """mocb200_options.persistent: every plain inner of a sweep call in one cooperative launch (grid barriers between
    the boundary phases, flux / q-bar update inside the kernel) gives what the per-phase launches give: flux,
    boundary flux and the coarse tallies of the last inner, Gauss-Seidel and Jacobi boundary update, one and two
    groups per call."""
    flat, gold = load_case(case)
    G, n_reg = (int(flat[k][0]) for k in ("n_group", "n_reg"))
    bcpg = int(flat["bc_per_group"][0])
    rng = np.random.default_rng(11)
    xstr = np.stack([gold[f"xs_tr_{g}"] for g in range(G)])
    xself = np.stack([gold[f"xs_self_{g}"] for g in range(G)])
    src = rng.uniform(0.05, 1.0, size=(G, n_reg))
    flux0 = rng.uniform(0.5, 1.5, size=(G, n_reg))
    bc = rng.uniform(0.0, 0.3, size=(G, bcpg))
    res = []
    for mode in (0, persistent):
        sw = _sweeper(flat, boundary_update=1 if jacobi else 0, kernel=5, persistent=mode)
        sw.set_xs(0, xstr, xstr_src=xstr, xs_self=xself)
        sw.set_source(0, src)
        sw.set_flux(0, flux0)
        sw.set_boundary(0, 0, bc)
        launches0 = sw.stats()["sweep_launches"]
        sw.sweep(0, 1, n_inner=n_inner, tally_mode=tally)
        sw.sweep(1, 2, n_inner=n_inner, tally_mode=tally)  # two groups per call
        n_launch = sw.stats()["sweep_launches"] - launches0
        out = [sw.get_flux(0, 3)] + [sw.get_boundary(0, g, 1)[0] for g in range(3)]
        if tally:
            for g in range(3):
                out.extend(sw.get_coarse(g))
        res.append((out, n_launch))
        sw.close()
    phases = 1 if jacobi else 2
    plain = n_inner - (1 if tally else 0)
    assert res[0][1] == 2 * phases * n_inner
    assert res[1][1] == 2 * ((1 if plain else 0) + (phases if tally else 0)), "the persistent path was not taken"
    for x, y in zip(res[0][0], res[1][0]):
        _close(x, y, atol=1e-13)


@pytest.mark.parametrize("case", ["mini2d_gs", "mini3d_gs"])
@pytest.mark.parametrize("slots", [1, 2])
def test_sliding_attenuation_cache_equals_full_cache(case, slots):
    """cache_groups < G (what a problem too large for the device gets): the cache is rebuilt for the groups of every
    call; results equal those with all groups resident. Group-batched sweeps are refused."""
    flat, gold = load_case(case)
    G, n_reg, n_plane = (int(flat[k][0]) for k in ("n_group", "n_reg", "n_plane"))
    bcpg = int(flat["bc_per_group"][0])
    rng = np.random.default_rng(17)
    xstr = np.stack([gold[f"xs_tr_{g}"] for g in range(G)])
    xself = np.stack([gold[f"xs_self_{g}"] for g in range(G)])
    src = rng.uniform(0.05, 1.0, size=(G, n_reg))
    bc = rng.uniform(0.0, 0.3, size=(n_plane, G, bcpg))
    res = []
    for cg in (0, slots):
        sw = _sweeper(flat, boundary_update=0, cache_groups=cg)
        sw.set_xs(0, xstr, xstr_src=xstr, xs_self=xself)
        sw.set_source(0, src)
        sw.set_flux(0, np.ones((G, n_reg)))
        for ip in range(n_plane):
            sw.set_boundary(ip, 0, bc[ip])
        for _ in range(2):  # two outers: every group's slot is evicted and rebuilt
            for g in range(G):
                sw.sweep(g, 1, n_inner=2, tally_mode=1)
        if slots >= 2:
            sw.sweep(0, 2, n_inner=1, tally_mode=0)  # two groups per call fit two slots
        elif cg:
            with pytest.raises(RuntimeError, match="attenuation cache holds"):
                sw.sweep(0, 2, n_inner=1, tally_mode=0)
        if cg:
            with pytest.raises(RuntimeError, match="group-batched"):
                sw.sweep(0, G, n_inner=1, tally_mode=0)
        out = [sw.get_flux(0, G)]
        out += [np.stack([sw.get_boundary(ip, g, 1)[0] for ip in range(n_plane)]) for g in range(G)]
        res.append(out)
        sw.close()
    for x, y in zip(*res):
        _close(x, y, atol=1e-13)


@pytest.mark.parametrize("case,parts", [("mini2d_gs", 2), ("3x3_s05_gs", 3)])
@pytest.mark.parametrize("tally", [0, 1])
def test_angle_family_sharding_equals_whole_sweep(case, parts, tally):
    """One plane split over `parts` handles by angle families (mocb200_options.family_begin/end): every handle sweeps
    its families (mocb200_sweep_partial), the FSR tallies and the coarse tallies are summed over the handles -- here
    on one GPU through adopted torch tensors, across ranks bench.py --shard angles all-reduces the same buffers over
    NCCL --, then mocb200_finalize_flux. Three inner iterations, Gauss-Seidel boundary update: flux, coarse tallies and
    the boundary flux of every family equal the single-handle sweep."""
    import torch
    from mocc_b200.capi import BUF_CURRENT, BUF_SURFACE_FLUX, BUF_TALLY, angle_families
    from mocc_b200.sharding import partition_families
    flat, gold = load_case(case)
    G, n_reg, bcpg = (int(flat[k][0]) for k in ("n_group", "n_reg", "bc_per_group"))
    n_fam, fam = angle_families(flat)
    ranges = partition_families(flat, fam, parts)
    assert len(ranges) == parts
    g, n_inner = 1, 3
    rng = np.random.default_rng(3)
    xstr, xself = gold[f"xs_tr_{g}"], gold[f"xs_self_{g}"]
    src = rng.uniform(0.05, 1.0, size=n_reg)
    flux0 = rng.uniform(0.5, 1.5, size=n_reg)
    bc = rng.uniform(0.0, 0.3, size=bcpg)

    def setup(**kw):
        sw = _sweeper(flat, boundary_update=0, **kw)
        sw.set_xs(g, xstr, xstr_src=xstr, xs_self=xself)
        sw.set_source(g, src)
        sw.set_flux(g, flux0)
        sw.set_boundary(0, g, bc)
        return sw
    whole = setup()
    whole.sweep(g, 1, n_inner=n_inner, tally_mode=tally)
    ref_flux, ref_bc = whole.get_flux(g, 1)[0], whole.get_boundary(0, g, 1)[0]
    ref_coarse = whole.get_coarse(g) if tally else None
    with pytest.raises(RuntimeError):  # a partial handle refuses the whole-sweep entry point
        s = setup(family_begin=ranges[0][0], family_end=ranges[0][1])
        try:
            s.sweep(g, 1)
        finally:
            s.close()
    whole.close()

    handles = [setup(family_begin=b, family_end=e) for b, e in ranges]
    bufs = []
    for sw in handles:
        mine = {}
        for which in (BUF_TALLY, BUF_CURRENT, BUF_SURFACE_FLUX):
            _, n = sw.device_buffer(which)
            t = torch.zeros(n, dtype=torch.float64, device="cuda")
            sw.adopt_device_buffer(which, t.data_ptr(), n)
            mine[which] = t
        bufs.append(mine)

    def reduce_over_handles(which):
        for sw in handles:
            sw.synchronize()
        total = sum(b[which] for b in bufs)
        for b in bufs:
            b[which].copy_(total)
        torch.cuda.synchronize()
    for inner in range(n_inner):
        last = inner == n_inner - 1
        for sw in handles:
            sw.sweep_partial(g, 1, tally_mode=tally if last else 0)
        reduce_over_handles(BUF_TALLY)
        if last and tally:
            reduce_over_handles(BUF_CURRENT)
            reduce_over_handles(BUF_SURFACE_FLUX)
        for sw in handles:
            sw.finalize_flux(g, 1)
    # boundary slots of a family belong to the handle that sweeps it
    off, sx, sy = flat["bc_offset"], flat["bc_size_x"], flat["bc_size_y"]
    got_bc = np.full(bcpg, np.nan)
    for sw, (b, e) in zip(handles, ranges):
        _close(sw.get_flux(g, 1)[0], ref_flux, rtol=1e-11)
        if tally:
            for x, y in zip(sw.get_coarse(g), ref_coarse):
                _close(x, y, rtol=1e-10, atol=1e-12)
        mine = sw.get_boundary(0, g, 1)[0]
        for ao in range(len(fam)):
            if b <= fam[ao] < e:
                sl = slice(int(off[ao]), int(off[ao] + sx[ao] + sy[ao]))
                got_bc[sl] = mine[sl]
        sw.close()
    assert not np.isnan(got_bc).any()
    _close(got_bc, ref_bc, rtol=1e-11)


def _source_tables(gold, G):
    from mocc_b200.capi import material_tables
    xs_nf = np.stack([gold[f"xs_nf_{g}"] for g in range(G)])
    xs_ch = np.stack([gold[f"xs_ch_{g}"] for g in range(G)])
    xs_scat = np.stack([gold[f"xs_scat_to_{g}"].reshape(G, -1) for g in range(G)])  # [to][from][n_reg]
    return xs_nf, xs_ch, xs_scat, material_tables(xs_nf, xs_ch, xs_scat)


@pytest.mark.parametrize("case", ["mini2d_gs", "mini3d_gs", "3x3_s05_gs"])
@pytest.mark.parametrize("external", [False, True])
def test_device_source_construction_is_bit_identical_to_the_oracle(case, external):
    """mocb200_fission_source / mocb200_build_source (calc_fission_source, Source::fission + in_scatter on the device)
    against the oracle's restatement, which is pinned bit for bit on the reference's own sources: same bits."""
    from oracle_lib import oracle_fission_source, oracle_group_source
    flat, gold = load_case(case)
    G, n_reg = (int(flat[k][0]) for k in ("n_group", "n_reg"))
    xs_nf, xs_ch, xs_scat, (fsr_mat, t_nf, t_ch, t_scat) = _source_tables(gold, G)
    assert t_nf.shape[0] < 16  # a handful of cross-section regions, not one per FSR
    rng = np.random.default_rng(17)
    flux = rng.uniform(0.2, 2.0, size=(G, n_reg))
    ext = rng.uniform(0.0, 0.1, size=(G, n_reg)) if external else None
    k = 1.0437
    sw = _sweeper(flat)
    with pytest.raises(RuntimeError):
        sw.fission_source(k)  # no tables yet
    sw.set_source_xs(fsr_mat, t_nf, t_ch, t_scat)
    with pytest.raises(RuntimeError):
        sw.build_source(0, 1)  # no fission source yet
    sw.set_flux(0, flux)
    sw.set_external_source(ext)
    sw.fission_source(k)
    fs = sw.get_fission_source()
    fs_o = oracle_fission_source(k, xs_nf, flux.T)
    assert np.array_equal(fs, fs_o)
    sw.build_source(0, G)           # all groups at once: every group sees the same flux
    src = sw.get_source(0, G)
    for g in range(G):
        s_o = oracle_group_source(g, xs_ch[g], fs_o, xs_scat[g], flux.T, ext=None if ext is None else ext[g])
        assert np.array_equal(src[g], s_o)
    # the host's fission source instead of the device's
    sw.set_fission_source(2.0 * fs_o)
    sw.build_source(1, 1)
    s_o = oracle_group_source(1, xs_ch[1], 2.0 * fs_o, xs_scat[1], flux.T, ext=None if ext is None else ext[1])
    assert np.array_equal(sw.get_source(1, 1)[0], s_o)
    sw.close()


@pytest.mark.parametrize("case", ["mini2d_gs", "3x3_s05_gs"])
def test_fixed_source_step_with_device_sources_equals_host_sources(case):
    """One FixedSourceSolver::step (fixed_source_solver.cpp:102-117) -- Gauss-Seidel over the groups, every group's
    source built from the fluxes swept so far -- with the sources built on the device from the resident flux, against
    the same step with the sources built on the host (oracle) and uploaded group by group."""
    from oracle_lib import oracle_fission_source, oracle_group_source
    flat, gold = load_case(case)
    G, n_reg, bcpg = (int(flat[k][0]) for k in ("n_group", "n_reg", "bc_per_group"))
    xs_nf, xs_ch, xs_scat, (fsr_mat, t_nf, t_ch, t_scat) = _source_tables(gold, G)
    xstr = np.stack([gold[f"xs_tr_{g}"] for g in range(G)])
    xself = np.stack([gold[f"xs_self_{g}"] for g in range(G)])
    rng = np.random.default_rng(23)
    flux0 = rng.uniform(0.5, 1.5, size=(G, n_reg))
    bc = np.full((G, bcpg), 1.0 / (4.0 * np.pi))
    k, n_inner = 0.97, 2
    res = []
    for device_sources in (True, False):
        sw = _sweeper(flat, boundary_update=0)
        sw.set_xs(0, xstr, xstr_src=xstr, xs_self=xself)
        sw.set_flux(0, flux0)
        sw.set_boundary(0, 0, bc)
        if device_sources:
            sw.set_source_xs(fsr_mat, t_nf, t_ch, t_scat)
            sw.fission_source(k)
            for g in range(G):
                sw.build_source(g, 1)
                sw.sweep(g, 1, n_inner=n_inner, tally_mode=1)
        else:
            fs = oracle_fission_source(k, xs_nf, flux0.T)
            for g in range(G):
                flux_now = sw.get_flux(0, G)
                sw.set_source(g, oracle_group_source(g, xs_ch[g], fs, xs_scat[g], flux_now.T))
                sw.sweep(g, 1, n_inner=n_inner, tally_mode=1)
        res.append([sw.get_flux(0, G)] + [sw.get_boundary(0, g, 1)[0] for g in range(G)] +
                   [x for g in range(G) for x in sw.get_coarse(g)])
        sw.close()
    for x, y in zip(*res):
        _close(x, y, rtol=1e-12, atol=1e-14)
