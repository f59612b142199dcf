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
