"""Sharding of the MoC sweep across GPUs / ranks (host-side logic): macroplanes, and angle families of one plane.

Inside a sweep the macroplanes are independent (reference: src/sweepers/moc/moc_sweeper_kernel.inc.hpp:51-153
loops planes outermost and every plane touches only its own boundary condition, FSR range and coarse-surface
range), so ranks own contiguous plane ranges and there is NO data-path collective. What the host solver needs
afterwards -- every plane's scalar flux and coarse tallies -- is assembled by an all-gather of the per-rank
slices (NCCL on GPUs, gloo in the CPU tests). Same partition rule as the C++ plugin
(mocc_b200/host/cuda_moc_sweeper.cpp).
"""
import numpy as np


def partition_planes(weights, n_parts):
    """Contiguous ranges [(begin, end), ...] balancing the summed weights (segments per plane)."""
    w = np.asarray(weights, dtype=np.float64)
    n = w.size
    n_parts = min(n_parts, n)
    total = w.sum()
    out, ip, done = [], 0, 0.0
    for k in range(n_parts):
        begin = ip
        target = total * (k + 1) / n_parts
        while ip < n and (ip == begin or done + 0.5 * w[ip] < target) and n - ip > n_parts - 1 - k:
            done += w[ip]
            ip += 1
        if k == n_parts - 1:
            ip = n
        out.append((begin, ip))
    return out


def plane_weights(arrays):
    """Segments swept per macroplane (all sweep angles), from a flattened problem."""
    n_geom = int(arrays["n_geom"][0])
    gtb, tsb = arrays["geom_trk_begin"], arrays["trk_seg_begin"]
    ang_geom = arrays["ang_geom"]
    w = []
    for u in arrays["plane_unique"]:
        per_geom = [tsb[gtb[u * n_geom + g + 1]] - tsb[gtb[u * n_geom + g]] for g in range(n_geom)]
        w.append(float(sum(per_geom[g] for g in ang_geom)))
    return w


def reg_range(arrays, plane_range):
    first = list(arrays["plane_first_reg"]) + [int(arrays["n_reg"][0])]
    return first[plane_range[0]], first[plane_range[1]]


def all_gather_flux(local_flux, arrays, ranges, rank, dist_module):
    """Assemble the full [..., n_reg] flux from per-rank arrays that are only valid on the rank's own planes."""
    import torch
    world = len(ranges)
    lo, hi = reg_range(arrays, ranges[rank])
    n_max = max(reg_range(arrays, r)[1] - reg_range(arrays, r)[0] for r in ranges)
    lead = local_flux.shape[:-1]
    send = torch.zeros(lead + (n_max,), dtype=torch.float64)
    send[..., : hi - lo] = torch.as_tensor(np.ascontiguousarray(local_flux[..., lo:hi]))
    recv = [torch.zeros_like(send) for _ in range(world)]
    dist_module.all_gather(recv, send)
    full = np.zeros(lead + (int(arrays["n_reg"][0]),))
    for r, rng in enumerate(ranges):
        a, b = reg_range(arrays, rng)
        full[..., a:b] = recv[r][..., : b - a].numpy()
    return full


# ---- angle families: one 2-D plane over several ranks ("for single-plane 2-D cases") ----
# A family (mocb200_angle_families) is closed under track reversal, polar bundling and the boundary update
# (boundary_condition.cpp:155-191), so a rank sweeps its families exactly as the whole sweep does, Gauss-Seidel
# order included, and NO boundary flux travels. What the reference sums over its threads -- the per-FSR tally t_flux,
# moc_sweeper_kernel.inc.hpp:155-163 -- is summed over the ranks (all-reduce), then every rank applies the flux
# update (:165-173). The coarse tallies of the last inner are summed the same way.

def family_weights(arrays, family):
    """Segments swept per angle family (forward angles only: a track is swept in both directions)."""
    n_ang = int(arrays["n_ang"][0])
    gtb, tsb, ang_geom = arrays["geom_trk_begin"], arrays["trk_seg_begin"], arrays["ang_geom"]
    w = np.zeros(int(np.max(family)) + 1)
    for a in range(n_ang):
        g = int(ang_geom[a])
        w[family[a]] += float(tsb[gtb[g + 1]] - tsb[gtb[g]])
    return w


def partition_families(arrays, family, n_parts):
    """Contiguous family ranges [(begin, end), ...] balancing the segments swept."""
    return partition_planes(family_weights(arrays, family), n_parts)


def allreduce_sum(x, dist_module):
    """Sum of a float64 array over the ranks (gloo on CPU; on GPUs the device buffers the handle adopted are
    all-reduced in place over NCCL, bench.py --shard angles)."""
    import torch
    t = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).clone()
    dist_module.all_reduce(t)
    return t.numpy()
