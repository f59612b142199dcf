"""Host-side multi-rank logic on CPU: plane partition and the all-gather of per-rank plane slices, exercised with
world_size 2 over gloo. The per-rank compute is the CPU oracle (test infrastructure) restricted to the rank's own
macroplanes -- on GPUs every rank runs the CUDA sweep on its planes instead (mocb200_options.plane_begin/end)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_case, records


def test_partition_is_contiguous_balanced_and_complete():
    from mocc_b200.sharding import partition_planes
    for w, n in (([1.0] * 8, 3), ([5, 1, 1, 1, 5, 1, 1, 1], 4), ([3, 3], 8), ([1, 2, 3, 4, 5, 6, 7], 2)):
        parts = partition_planes(w, n)
        assert parts[0][0] == 0 and parts[-1][1] == len(w)
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all(e > b for b, e in parts)
        assert len(parts) == min(n, len(w))
    assert partition_planes([1.0] * 8, 4) == [(0, 2), (2, 4), (4, 6), (6, 8)]
    loads = [sum([5, 1, 1, 1, 5, 1, 1, 1][b:e]) for b, e in partition_planes([5, 1, 1, 1, 5, 1, 1, 1], 2)]
    assert max(loads) <= 8


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from mocc_b200.sharding import all_gather_flux, partition_planes, plane_weights, reg_range
    from oracle_lib import oracle_sweep1g
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat, gold = load_case("mini3d_gs")
    rec = records(gold)[0]
    ranges = partition_planes(plane_weights(flat), world)
    # the rank's own planes: the oracle sweeps every plane, a rank keeps only its FSR range
    flux, _, _, _ = oracle_sweep1g(flat, rec["xstr"], rec["qbar"], rec["bc_in"], gs_boundary=True)
    lo, hi = reg_range(flat, ranges[rank])
    local = np.full_like(flux, np.nan)
    local[lo:hi] = flux[lo:hi]
    full = all_gather_flux(local, flat, ranges, rank, dist)
    np.save(os.path.join(out_dir, f"full_{rank}.npy"), full)
    np.save(os.path.join(out_dir, f"ranges_{rank}.npy"), np.array(ranges))
    dist.destroy_process_group()


def test_two_ranks_gather_plane_slices(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    flat, gold = load_case("mini3d_gs")
    rec = records(gold)[0]
    ranges = np.load(tmp_path / "ranges_0.npy")
    assert ranges.shape == (2, 2) and ranges[0, 0] == 0 and ranges[1, 1] == int(flat["n_plane"][0])
    for r in range(2):
        full = np.load(tmp_path / f"full_{r}.npy")
        assert not np.isnan(full).any()
        assert np.array_equal(full, rec["flux_out"])  # the oracle is bit-identical to the reference record


def test_stacked_planes_sweep_like_the_single_plane():
    """bench.py --gpus N sweeps ONE stack of N copies of the workload's plane (rank r owns plane r): the stacked
    problem must be a valid multi-plane problem whose planes behave exactly like the original one. Checked with
    the CPU oracle: every plane of the 3-plane stack reproduces the single-plane sweep bit for bit (flux, boundary
    flux, coarse tallies at the plane's own surface range)."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle_lib import oracle_sweep1g
    flat, gold = load_case("mini2d_gs")
    rec = records(gold)[1]
    mode = int(rec["mode"][0])
    f1, bc1, cur1, sf1 = oracle_sweep1g(flat, rec["xstr"], rec["qbar"], rec["bc_in"], gs_boundary=True, tally_mode=mode)
    n = 3
    st = bench.stack_planes({k: v for k, v in flat.items()} | {"xs_tr": np.zeros((1, int(flat["n_reg"][0]))),
                                                               "xs_self": np.zeros((1, int(flat["n_reg"][0]))),
                                                               "xs_nf": np.zeros((1, int(flat["n_reg"][0]))),
                                                               "xs_ch": np.zeros((1, int(flat["n_reg"][0]))),
                                                               "xs_scat": np.zeros((1, 1, int(flat["n_reg"][0])))}, n)
    R, nsp = int(flat["n_reg"][0]), int(flat["n_surf_plane"][0])
    assert int(st["n_reg"][0]) == n * R and int(st["n_surf"][0]) == n * nsp + int(flat["n_cell_plane"][0])
    fN, bcN, curN, sfN = oracle_sweep1g(st, np.tile(rec["xstr"], n), np.tile(rec["qbar"], n),
                                        np.tile(rec["bc_in"].reshape(1, -1), (n, 1)), gs_boundary=True, tally_mode=mode)
    nxy = int(flat["nx"][0]) * int(flat["ny"][0])
    for ip in range(n):
        assert np.array_equal(fN[ip * R:(ip + 1) * R], f1)
        assert np.array_equal(bcN[ip], bc1[0])
        if mode == 1:  # radial surfaces of the plane (the first nx*ny of every plane's range are its bottom faces)
            assert np.array_equal(curN[ip * nsp + nxy:(ip + 1) * nsp], cur1[nxy:nsp])
            assert np.array_equal(sfN[ip * nsp + nxy:(ip + 1) * nsp], sf1[nxy:nsp])


# ---- angle families (one plane over several ranks) ----

def _families_py(flat):
    """Independent restatement of the closure the C ABI computes: union-find over track reversal, polar copies of an
    azimuth and the boundary-update destinations."""
    n_ang = int(flat["n_ang"][0])
    nab, ndo = 2 * n_ang, int(flat["ndir_oct"][0])
    parent = list(range(nab))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    def unite(a, b):
        a, b = find(a), find(b)
        if a != b:
            parent[max(a, b)] = min(a, b)
    off, sx, sy = flat["bc_offset"], flat["bc_size_x"], flat["bc_size_y"]
    for a in range(n_ang):
        unite(a, a + n_ang)
    for o in range(2):
        for a in range(o * ndo, (o + 1) * ndo):
            for b in range(a + 1, (o + 1) * ndo):
                if abs(flat["ang_alpha"][a] - flat["ang_alpha"][b]) < 1e-12:
                    unite(a, b)
    for ao in range(nab):
        for face in range(2):
            if flat["bc_dst_kind"][2 * ao + face] == 2:
                continue
            dst = int(flat["bc_dst_off"][2 * ao + face])
            t = next(t for t in range(nab) if off[t] <= dst < off[t] + sx[t] + sy[t])
            unite(ao, t)
    roots = sorted({find(i) for i in range(nab)})
    return [roots.index(find(i)) for i in range(nab)]


@pytest.mark.parametrize("case", ["mini2d_gs", "3x3_s05_gs", "mini3d_gs"])
def test_angle_families_are_closed_under_reversal_and_boundary_update(case):
    from mocc_b200.capi import angle_families
    flat, _ = load_case(case)
    n, fam = angle_families(flat)
    assert list(fam) == _families_py(flat)
    assert n == max(fam) + 1 and n > 1
    # every family holds angles of both boundary phases (octants 1 and 2): the Gauss-Seidel order survives
    n_ang, ndo = int(flat["n_ang"][0]), int(flat["ndir_oct"][0])
    for f in range(n):
        members = [a for a in range(n_ang) if fam[a] == f]
        assert any(a < ndo for a in members) and any(a >= ndo for a in members)


def _family_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from mocc_b200.capi import angle_families
    from mocc_b200.sharding import allreduce_sum, partition_families
    from oracle_lib import oracle_sweep1g
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat, gold = load_case("mini2d_gs")
    rec = records(gold)[0]
    n_ang = int(flat["n_ang"][0])
    n, fam = angle_families(flat)
    lo, hi = partition_families(flat, fam, world)[rank]
    # the rank's share of the sweep: the oracle (test infrastructure) with the tally weights of the other families'
    # angles set to zero leaves exactly this rank's partial t_flux behind
    mine = dict(flat)
    mine["wt_v_st"] = np.where([lo <= fam[a] < hi for a in range(n_ang)], flat["wt_v_st"], 0.0)
    flux_part, bc_out, _, _ = oracle_sweep1g(mine, rec["xstr"], rec["qbar"], rec["bc_in"], gs_boundary=True)
    fpi = 4.0 * np.pi
    partial = (flux_part - rec["qbar"] * fpi) * (rec["xstr"] * flat["vol"])   # t_flux of my angles
    tally = allreduce_sum(partial, dist)                                          # kernel:155-163 across ranks
    flux = tally / (rec["xstr"] * flat["vol"]) + rec["qbar"] * fpi                # kernel:165-173
    np.save(os.path.join(out_dir, f"aflux_{rank}.npy"), flux)
    np.save(os.path.join(out_dir, f"abc_{rank}.npy"), bc_out)
    dist.destroy_process_group()


def test_two_ranks_sum_angle_family_tallies(tmp_path):
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_family_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    flat, gold = load_case("mini2d_gs")
    rec = records(gold)[0]
    for r in range(2):
        flux = np.load(tmp_path / f"aflux_{r}.npy")
        assert np.max(np.abs(flux - rec["flux_out"]) / rec["flux_out"]) < 1e-12
        assert np.array_equal(np.load(tmp_path / f"abc_{r}.npy").reshape(-1), rec["bc_out"].reshape(-1))
