"""The CPU oracle (oracle/moc_oracle.c) is pinned against the UNMODIFIED reference.

Golden records under tests/golden/ were produced by the reference's own sweep1g
(oracle/ref_tool.cpp, tests/golden/make_golden.sh, single-threaded). The oracle
must reproduce them: boundary flux and q-bar bit for bit; scalar flux and coarse
tallies bit for bit too (same summation order as the single-threaded reference).
Known answers of the reference's own unit tests are checked as well.
"""
import numpy as np
import pytest

from conftest import CASES, load_case, records
from oracle_lib import oracle_exp, oracle_self_scatter, oracle_sweep1g, oracle_sweep1g_corrections


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_sweep_matches_reference(case):
    flat, gold = load_case(case)
    gs = bool(gold["gs_boundary"][0])
    recs = records(gold)
    assert recs
    for rec in recs:
        mode = int(rec["mode"][0])
        if mode == 2:
            continue  # test_oracle_corrections_match_reference
        flux, bc, cur, sf = oracle_sweep1g(flat, rec["xstr"], rec["qbar"], rec["bc_in"], gs_boundary=gs,
                                           tally_mode=mode)
        assert np.array_equal(bc.ravel(), rec["bc_out"]), "boundary flux must be bit-identical"
        assert np.array_equal(flux, rec["flux_out"]), "scalar flux must be bit-identical (1-thread reference)"
        if mode == 1:
            assert np.array_equal(cur, rec["current"])
            assert np.array_equal(sf, rec["surface_flux"])


def test_oracle_corrections_match_reference():
    """cmdo::CurrentCorrections restated: currents, surface flux and the alpha/beta correction factors of the
    reference's MoCSweeper_2D3D (self-coupled) are reproduced bit for bit."""
    flat, gold = load_case("mini3d_2d3d")
    recs = [r for r in records(gold) if int(r["mode"][0]) == 2]
    assert len(recs) == 3
    for rec in recs:
        flux, bc, cur, sf, alpha, beta = oracle_sweep1g_corrections(
            flat, rec["xstr"], rec["xstr_true"], rec["qbar"], rec["sn_xs"], rec["bc_in"], gs_boundary=True)
        assert np.array_equal(bc.ravel(), rec["bc_out"])
        assert np.array_equal(flux, rec["flux_out"])
        assert np.array_equal(cur, rec["current"])
        assert np.array_equal(sf, rec["surface_flux"])
        assert np.array_equal(alpha.ravel(), rec["alpha"])
        assert np.array_equal(beta.ravel(), rec["beta"])


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_self_scatter_matches_reference(case):
    flat, gold = load_case(case)
    for rec in records(gold):
        g = int(rec["group"][0])
        q = oracle_self_scatter(rec["src"], rec["flux_in"], gold[f"xs_self_{g}"], gold[f"xs_tr_{g}"])
        assert np.array_equal(q, rec["qbar"])


def test_exponential_known_answers():
    # src/core/tests/test_Exponential.cpp:27-43: |table - exp| < 2e-8 on x = -10, -9.9, ...
    tab = np.array([np.exp(-10.0 + i * ((0.0 - -10.0) / 10000.0)) for i in range(10001)])
    tab = np.append(tab, tab[-1])
    x = -10.0
    while x < 0.0:
        assert abs(oracle_exp(tab, x) - np.exp(x)) < 2e-8
        x += 0.1
    # relative interpolation error of the 10000-interval table (SURVEY: 1.25e-7)
    xs = np.linspace(-10.0, -1e-9, 4001)
    assert max(abs(oracle_exp(tab, v) - np.exp(v)) / np.exp(v) for v in xs) < 1.3e-7
    # :66-91, Exponential_Linear<5>(-5.3, 0.0): data points and two interpolated known answers
    space = (0.0 - -5.3) / 5.0
    t5 = np.array([np.exp(-5.3 + i * space) for i in range(6)])
    t5 = np.append(t5, t5[-1])
    for i, xv in enumerate((-5.3, -4.24, -3.18, -2.12, -1.06)):
        assert oracle_exp(t5, xv, n=5, vmin=-5.3) == pytest.approx(t5[i], abs=1e-12)
    assert oracle_exp(t5, -5.088, n=5, vmin=-5.3) == pytest.approx(6.87479349415065e-03, abs=1e-12)
    assert oracle_exp(t5, -2.756, n=5, vmin=-5.3) == pytest.approx(7.29640444772866e-02, abs=1e-12)


def test_exp_table_in_flat_file_is_reference_table():
    flat, _ = load_case("mini2d_gs")
    tab = flat["exp_table"]
    assert tab.size == 10002
    space = (0.0 - -10.0) / 10000.0
    ref = np.array([np.exp(-10.0 + i * space) for i in range(10001)])
    # numpy's exp and glibc's std::exp may differ in the last ulp on some inputs
    assert np.max(np.abs(tab[:10001] - ref) / ref) < 3e-16
    assert tab[10001] == tab[10000]


@pytest.mark.parametrize("case", ["mini2d_gs", "mini2d_nocmfd", "mini3d_gs", "mini3d_2d3d", "3x3_s05_gs"])
def test_oracle_source_construction_is_bit_identical_to_the_reference(case):
    """Fission source (TransportSweeper::calc_fission_source) and the 1-group sources (Source::fission + in_scatter)
    the reference built while the goldens were recorded, reproduced bit for bit from the recorded fluxes."""
    from oracle_lib import oracle_fission_source, oracle_group_source
    flat, gold = load_case(case)
    G = int(flat["n_group"][0])
    xs_nf = np.stack([gold[f"xs_nf_{g}"] for g in range(G)])
    n_outer = int(gold["n_outer"][0])
    fs_of = {}
    for o in range(n_outer):
        fs = oracle_fission_source(float(gold[f"outer{o}_k"][0]), xs_nf, gold[f"outer{o}_flux_start"])
        assert np.array_equal(fs, gold[f"outer{o}_fission_source"])
        fs_of[o] = fs
    assert any(np.any(fs > 0) for fs in fs_of.values())
    for rec in records(gold):
        g, o = int(rec["group"][0]), int(rec["outer"][0])
        src = oracle_group_source(g, gold[f"xs_ch_{g}"], fs_of[o], gold[f"xs_scat_to_{g}"].reshape(G, -1), rec["flux_all"])
        assert np.array_equal(src, rec["src"])


def test_oracle_reaches_the_analytic_infinite_medium_spectrum():
    """The reference's own known-answer test of the sweep (src/sweepers/moc/tests/test_MoC_IHM.cpp:99-147), run on
    the ORACLE: infinite homogeneous medium (UO2-3.3, all boundaries reflective, LS-2, spacing 0.01), flux set to the
    analytic spectrum phi = M^-1 chi (M = diag(Sigma_tr) - Sigma_s, :160-181), fission source = 1 in every region, then
    group by group the reference's source construction and 800 inner iterations of self scatter + sweep: every FSR flux
    within 0.5 % of phi_g, the reference's tolerance (:141-144). The same test runs on the CUDA sweeper itself in
    tests/test_gpu_plugin.py (compiled from the reference's unmodified source)."""
    from oracle_lib import oracle_fission_source, oracle_group_source, oracle_self_scatter
    flat, gold = load_case("ihm")
    G, n_reg = int(flat["n_group"][0]), int(flat["n_reg"][0])
    assert G == 7 and n_reg == 54 and int(gold["n_inner"][0]) == 800
    tr = np.array([gold[f"xs_tr_{g}"][0] for g in range(G)])
    nf = np.array([gold[f"xs_nf_{g}"][0] for g in range(G)])
    chi = np.array([gold[f"xs_ch_{g}"][0] for g in range(G)])
    scat = np.array([gold[f"xs_scat_to_{g}"].reshape(G, -1)[:, 0] for g in range(G)])  # [to][from]
    phi = np.linalg.solve(np.diag(tr) - scat, chi)
    k_inf = float(nf @ phi)
    assert 0.5 < k_inf < 1.5 and np.all(phi > 0)  # 0.7382: what the reference solve of this input converges to as well
    xs_nf = np.stack([gold[f"xs_nf_{g}"] for g in range(G)])
    flux = np.tile(phi, (n_reg, 1))                      # [n_reg][G], the true spectrum (set_spectrum, :84-92)
    fs = oracle_fission_source(k_inf, xs_nf, flux)
    assert np.max(np.abs(fs - 1.0)) < 1e-14              # :127-129
    bc = np.zeros((G, int(flat["n_plane"][0]), int(flat["bc_per_group"][0])))
    for g in range(G):
        src = oracle_group_source(g, gold[f"xs_ch_{g}"], fs, gold[f"xs_scat_to_{g}"].reshape(G, -1), flux)
        f, b = flux[:, g].copy(), bc[g]
        for _ in range(800):
            q = oracle_self_scatter(src, f, gold[f"xs_self_{g}"], gold[f"xs_tr_{g}"])
            f, b, _, _ = oracle_sweep1g(flat, gold[f"xs_tr_{g}"], q, b, gs_boundary=True)
        flux[:, g] = f
        assert np.max(np.abs(f - phi[g])) < 0.005 * phi[g], f"group {g}: {np.max(np.abs(f - phi[g]) / phi[g]):.3e}"
