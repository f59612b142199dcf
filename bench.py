#!/usr/bin/env python
"""bench.py -- throughput of the MoC transport sweep on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode pergroup|batched] [--impl reference]
                    [--shard planes|angles] [--workload c5g7_2d|dense_a|dense_b|quarter_core] [--persistent 0|1|2]

Workload (config.workload): C5G7 2-D, `examples/c5g7_2d.xml` as shipped by the reference
(7 groups, Chebyshev-Gauss 8x2 per octant, ray spacing 0.05 cm, n_inner = 10, CMFD on so the
last inner of every group tallies coarse currents). One STEP is what
FixedSourceSolver::step asks of the sweeper (src/solvers/fixed_source_solver.cpp:102-117):
sweep(g) for every group g, each n_inner inner iterations = 2*S*G*n_inner segment-group
updates (S = reference segment count, polar copies counted).

  value   device-resident: sources/XS/boundary flux already in HBM, only sweeps are timed
  e2e     the same step through the C ABI with HOST buffers: per group the one-group source,
          scalar flux and boundary flux go host->device and flux, boundary flux and coarse
          tallies come back, exactly what the C++ plugin (mocc_b200/host) does per sweep(group)
  e2e_device_sources  the same step with the sources built ON THE DEVICE from the resident flux
          (mocb200_fission_source / mocb200_build_source): the host sends k, downloads as in e2e
  e2e_plugin, time_to_converge_s   MOCC's own eigenvalue solve of the same input through the C++ plugin
          (<sweeper type="moc_cuda">) next to the reference sweeper on the box's host cores
  mode    pergroup: one mocb200_sweep per group (the reference's sweep(group) contract,
          Gauss-Seidel in energy); batched: all groups in one mocb200_sweep (8 group lanes)

With N > 1 (torchrun, one rank per GPU) the problem is ONE stack of N C5G7 planes (N macroplanes of one
geometry, the plane sharding the 2D3D method uses); rank r creates its handle with plane_begin = r,
plane_end = r + 1. Planes are independent inside a sweep (no collective in `value`); the e2e leg adds what
the host solver needs after every sweep(group): the ranks' flux and coarse current / surface-flux slices,
packed on the device (mocb200_pack_results_device) and exchanged with ONE NCCL all-gather per sweep(group),
then copied to pinned host memory (everything on rank 0, where a single-process host solver would run; the own
planes on the other ranks); `comm` reports the device time of those all-gathers
(scaling "weak": one plane per GPU).
--shard angles (N > 1): ONE plane, rank r sweeps its angle families (mocb200_angle_families: closed under the boundary
update, so the Gauss-Seidel boundary order is kept and no boundary flux travels); after every inner sweep one NCCL
all-reduce sums the per-FSR tally on the device buffer the handle adopted, then every rank applies the flux update
(scaling "strong": the total work is fixed).

--impl reference times the UNMODIFIED reference CPU sweeper (oracle/_ref/ref_tool, OpenMP on
all host cores) on the same input; rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MoC segment-group updates/s per sweep"
UNIT = "updates/s"
BYTES_PER_UPDATE = 6.10  # SURVEY.md 8(d) contract figure, per-group sweep of C5G7-2D (reference layout)
BIN = os.path.join(ROOT, "mocc_b200", "bin")
REF_TOOL = os.path.join(ROOT, "oracle", "_ref", "ref_tool")
CACHE = os.environ.get("MOCC_B200_CACHE", "/tmp/mocc_b200_bench")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


WORKLOADS = {
    # BASELINE.json configs[1]: the reference's example as shipped (the headline workload)
    "c5g7_2d": ("C5G7 2-D (examples/c5g7_2d.xml): 7 groups, cg 8x2, spacing 0.05, n_inner 10, coarse-current tally on "
                "the last inner", []),
    # BASELINE.json configs[2] / SURVEY.md 8(d) "dense-A": same file, denser quadrature and rays (~x10 segments)
    "dense_a": ("C5G7 2-D dense-A (examples/c5g7_2d.xml with cg 16x4, spacing 0.02): 7 groups, n_inner 10, "
                "coarse-current tally on the last inner",
                ["solver/ang_quad@n_azimuthal=16", "solver/ang_quad@n_polar=4", "solver/sweeper/rays@spacing=0.02"]),
    # SURVEY.md 8(d) "dense-B" (~x40 segments: S ~ 7.3e8, the attenuation cache of all 7 groups is ~41 GB)
    "dense_b": ("C5G7 2-D dense-B (examples/c5g7_2d.xml with cg 32x4, spacing 0.01): 7 groups, n_inner 10, "
                "coarse-current tally on the last inner",
                ["solver/ang_quad@n_azimuthal=32", "solver/ang_quad@n_polar=4", "solver/sweeper/rays@spacing=0.01"]),
    # BASELINE.json configs[4] at plane level: 9x9 checkerboard of the C5G7 assemblies with a reflector ring
    # (tools/make_quarter_core.py); with --gpus N one such plane per GPU
    "quarter_core": ("synthetic quarter core, 9x9 assemblies of 17x17 pins (C5G7 lattices, cg 8x2, spacing 0.05): "
                     "7 groups, n_inner 10, coarse-current tally on the last inner", None),
}


def workload_files(name="c5g7_2d"):
    """Flatten examples/c5g7_2d.xml with the plugin's own setup code (mocc_flatten)."""
    os.makedirs(CACHE, exist_ok=True)
    flat = os.path.join(CACHE, f"{name}.mocflat")
    if not os.path.exists(flat):
        tool = os.path.join(BIN, "mocc_flatten")
        inputs = os.path.join(BIN, "inputs")
        if not os.path.exists(tool):
            raise RuntimeError("mocc_b200/bin/mocc_flatten missing: run __graft_entry__.build() where the "
                               "reference sources are available")
        tmp = flat + f".tmp{os.getpid()}"
        if WORKLOADS[name][1] is None:  # an authored input: generated next to the example's data files
            work = os.path.join(CACHE, name)
            os.makedirs(work, exist_ok=True)
            subprocess.check_call(["cp", os.path.join(inputs, "c5g7.xsl"), work])
            subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_quarter_core.py"),
                                   os.path.join(inputs, "c5g7_2d.xml"), os.path.join(work, "in.xml")],
                                  stdout=subprocess.DEVNULL)
            subprocess.check_call([tool, "in.xml", tmp, "--xs"], cwd=work, stdout=subprocess.DEVNULL)
        else:
            sets = [x for s in WORKLOADS[name][1] for x in ("--set", s)]
            subprocess.check_call([tool, "c5g7_2d.xml", tmp, "--xs"] + sets, cwd=inputs, stdout=subprocess.DEVNULL)
        os.replace(tmp, flat)
    return flat


L2_NOTE = ("working set per step (segment + attenuation streams, > 1 GB) exceeds the 126 MB L2 and every host cache; "
           "the B200 arm also writes a 256 MB flush buffer between timed steps")


def workload_config(name, mode, boundary, S, n_reg, G, n_inner, world):
    """`config` of the JSON line: the WORKLOAD only, identical for the B200 arm and the reference arm."""
    return {"workload": WORKLOADS[name][0] + (f"; {world} axial planes, one per GPU" if world > 1 else ""),
            "mode": mode, "boundary_update": boundary, "segments": int(S), "n_reg": int(n_reg), "groups": int(G),
            "n_inner": int(n_inner), "updates_per_step": 2.0 * S * G * n_inner * world, "l2": L2_NOTE}


def stack_planes(arr, n):
    """An n-plane problem out of a one-plane one: n macroplanes sharing the plane's ray data (unique geometry 0),
    FSRs / coarse cells / coarse surfaces numbered plane by plane as Mesh does (mesh.cpp:88-136)."""
    if n == 1:
        return arr
    assert int(arr["n_plane"][0]) == 1
    out = dict(arr)
    R, ncp, nsp = int(arr["n_reg"][0]), int(arr["n_cell_plane"][0]), int(arr["n_surf_plane"][0])
    i32 = lambda v: np.array([v], dtype=np.int32)  # noqa: E731
    out["n_plane"], out["nz"], out["n_reg"] = i32(n), i32(n), i32(n * R)
    out["n_cell"], out["n_surf"] = i32(n * ncp), i32(n * nsp + ncp)
    for k in ("wt_v_st", "cur_wx", "cur_wy", "flx_wx", "flx_wy", "plane_height", "plane_dz", "vol"):
        out[k] = np.tile(arr[k], n)
    out["plane_unique"] = np.zeros(n, dtype=np.int32)
    out["plane_first_reg"] = (np.arange(n) * R).astype(np.int32)
    out["plane_cell_offset"] = (np.arange(n) * ncp).astype(np.int32)
    out["plane_xs_offset"] = (np.arange(n) * ncp).astype(np.int32)
    out["plane_surf_offset"] = (np.arange(n) * nsp).astype(np.int32)
    out["surf_area"] = np.concatenate([np.tile(arr["surf_area"][:nsp], n), arr["surf_area"][nsp:]])
    out["n_seg_reference"] = arr["n_seg_reference"] * n
    out["n_ray_reference"] = arr["n_ray_reference"] * n
    for k in ("xs_tr", "xs_self", "xs_nf", "xs_ch"):
        out[k] = np.tile(arr[k], (1, n))
    out["xs_scat"] = np.tile(arr["xs_scat"], (1, 1, n))
    return out


def synthetic_source(arr, G, n_reg):
    """Fixed one-group sources from a flat unit flux: fission (k = 1) + in-scatter; synthetic."""
    nf, ch, scat = arr["xs_nf"], arr["xs_ch"], arr["xs_scat"]
    fis = nf.sum(axis=0)  # sum_g nu-fission * flux(=1)
    src = np.empty((G, n_reg))
    for g in range(G):
        inscat = scat[g].sum(axis=0) - scat[g, g]
        src[g] = ch[g] * fis + inscat
    return src


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons = [], 0.0, set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = max(smax, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        busy = [x for x in sm if x > 0.5 * smax] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """The reference's own CPU sweep (unmodified sources, oracle/_ref) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if not os.path.exists(REF_TOOL):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_tool not built"}))
        return
    # Same workload as the B200 arm: the XML's n_inner (10) is NEVER changed, so the share of tallying
    # (moc::Current) inners is the same 1 in 10. The run is bounded instead by sweeping only the first
    # `ng` of the 7 groups in every step when K + W full steps would exceed the CPU-time budget (every
    # group costs the same: same rays, same inners); updates are counted for the groups actually swept.
    total = args.steps + args.warmup
    G, n_inner = 7, 10
    est_full_step_s = 2.55e9 / (9.0e7 * min(cores, 16))  # ~9e7 updates/s per host thread (r1 measurement)
    budget_s = float(os.environ.get("MOCC_B200_REF_BUDGET_S", "150"))
    ng = int(max(1, min(G, budget_s / (total * est_full_step_s) * G)))
    inputs = os.path.join(ROOT, "oracle", "_ref", "inputs")
    env = dict(os.environ, OMP_NUM_THREADS=str(cores))
    cmd = [REF_TOOL, "time", "c5g7_2d.xml", "--cmfd", "--sweeps", str(args.steps), "--warmup", str(args.warmup),
           "--groups", str(ng)]
    out = subprocess.run(cmd, cwd=inputs, env=env, capture_output=True, text=True, check=True).stdout
    res = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
    assert res["n_inner"] == n_inner, "reference arm must run the XML's n_inner"
    value = res["updates_per_s"]
    sample = (f"{args.steps} timed steps (+{args.warmup} warm-up) of c5g7_2d.xml, each {res['groups']} of 7 groups x "
              f"{n_inner} inner sweeps (last inner with the moc::Current tally), {res['updates']:.3e} updates in "
              f"{res['seconds']:.2f} s")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["seconds"] / args.steps * 1e3 * G / res["groups"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "examples/c5g7_2d.xml geometry and cross sections, flat initial flux",
        # the same `config` object as the B200 arm prints for this workload (implementation details of either arm
        # live in `arm`, not in `config`)
        "config": dict(workload_config("c5g7_2d", "pergroup", "gs", res["segments"], res["n_reg"], G, n_inner, 1),
                       **({"groups_sampled_per_step": res["groups"]} if res["groups"] != G else {})),
        "arm": {"sweeper": "reference MoCSweeper on CPU (OpenMP), unmodified sources (oracle/_ref)",
                "threads": res["threads"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["threads"], "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_angle_sharded(args, arr, world, rank, local, barrier):
    """N ranks on ONE plane (strong scaling): rank r sweeps its angle families with the reference's
    Gauss-Seidel boundary order intact (a family is closed under the boundary update: no boundary flux travels); after
    every inner sweep the per-FSR tally is summed over the ranks by ONE ncclAllReduce on the device buffer the handle
    adopted, then every rank applies the flux update; the coarse tallies of the last inner are summed the same way."""
    import torch
    import torch.distributed as dist
    from mocc_b200 import Sweeper
    from mocc_b200.capi import BUF_CURRENT, BUF_SURFACE_FLUX, BUF_TALLY, TALLY_CURRENT, angle_families
    from mocc_b200.sharding import partition_families
    S = int(arr["n_seg_reference"][0])
    G, n_reg = int(arr["n_group"][0]), int(arr["n_reg"][0])
    bcpg, n_surf, n_inner = int(arr["bc_per_group"][0]), int(arr["n_surf"][0]), int(arr["n_inner"][0])
    n_fam, fam = angle_families(arr)
    ranges = partition_families(arr, fam, world)
    if len(ranges) < world:
        raise RuntimeError(f"{n_fam} angle families cannot be split over {world} ranks")
    fb, fe = ranges[rank]
    sw = Sweeper(arr, device=local, boundary_update=0, kernel=args.kernel, max_polar=args.max_polar,
                 family_begin=fb, family_end=fe)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sw.set_stream(stream.cuda_stream)
    src = synthetic_source(arr, G, n_reg)
    sw.set_xs(0, arr["xs_tr"], xstr_src=arr["xs_tr"], xs_self=arr["xs_self"])
    sw.set_source(0, src)
    sw.set_flux(0, np.ones((G, n_reg)))
    bc0 = np.full((G, bcpg), 1.0 / (4.0 * np.pi))
    sw.set_boundary(0, 0, bc0)
    bufs = {}
    for which in (BUF_TALLY, BUF_CURRENT, BUF_SURFACE_FLUX):
        _, n = sw.device_buffer(which)
        bufs[which] = torch.zeros(n, dtype=torch.float64, device="cuda")
        sw.adopt_device_buffer(which, bufs[which].data_ptr(), n)
    stride = bufs[BUF_TALLY].numel() // G       # one group of the group-major tally
    tally1 = bufs[BUF_TALLY][:stride]
    comm_ev = []

    def sweep_group(g, time_comm=False):
        for inner in range(n_inner):
            last = inner == n_inner - 1
            sw.sweep_partial(g, 1, tally_mode=TALLY_CURRENT if last else 0)
            if time_comm:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            dist.all_reduce(tally1)
            if last:
                dist.all_reduce(bufs[BUF_CURRENT])
                dist.all_reduce(bufs[BUF_SURFACE_FLUX])
            if time_comm:
                e1.record(stream)
                comm_ev.append((e0, e1))
            sw.finalize_flux(g, 1)

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()
    src_h, flux_h, bc_h = pinned(src), pinned(np.ones((G, n_reg))), pinned(bc0)

    def step_e2e():
        for g in range(G):
            sw.set_sweep_inputs(g, src_h[g], flux_h[g], [bc_h[g]])
            sweep_group(g)
            # every rank keeps flux and its own boundary flux; the coarse tallies go where the host solver runs
            sw.get_sweep_results(g, flux_h[g], [bc_h[g]], coarse=rank == 0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(args.warmup):
        for g in range(G):
            sweep_group(g)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = sw.stats()["kernel_launches"]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for e0, e1 in ev:
        flush.zero_()
        barrier()  # strong scaling: every step starts together (untimed)
        e0.record(stream)
        for g in range(G):
            sweep_group(g)
        e1.record(stream)
    barrier()
    ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    ln = torch.tensor([float(sw.stats()["kernel_launches"] - launches0)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(ln, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    updates_step = 2.0 * S * G * n_inner       # ONE plane, whatever N
    value = updates_step * args.steps / (ms * 1e-3)
    # sweep kernels of this rank's families, per inner (CUDA events inside the library)
    sw.set_timing(True)
    for g in range(G):
        sweep_group(g)
    k_ms, k_n = sw.get_timing()
    sw.set_timing(False)
    for g in range(G):
        sweep_group(g, time_comm=True)
    torch.cuda.synchronize()
    cm = torch.tensor([sum(a.elapsed_time(b) for a, b in comm_ev)], dtype=torch.float64, device="cuda")
    dist.all_reduce(cm, op=dist.ReduceOp.MAX)
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = updates_step * args.steps / float(te.item())
    clocks = sampler.stop() if sampler else None
    peak, peak_src = peaks()
    sweep_ms = k_ms / max(k_n, 1)
    bytes_contract = BYTES_PER_UPDATE * 2.0 * S / world if args.workload == "c5g7_2d" else None
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64",
            "data": "examples/c5g7_2d.xml geometry and cross sections (flattened on the box); synthetic fixed source",
            "config": dict(workload_config(args.workload, "pergroup", "gs", S, n_reg, G, n_inner, 1),
                           shard=f"angle families: {n_fam} families over {world} ranks {ranges}, Gauss-Seidel boundary "
                                 "order kept, no boundary exchange"),
            "arm": {"kernel": "rchunk", "families": [fb, fe]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * G * 8 * (2 * n_reg + bcpg),
                    "d2h_bytes_per_step": G * 8 * (world * (n_reg + bcpg) + 2 * n_surf)},
            "comm": {"collective": "ncclAllReduce of the per-FSR tally after every inner sweep (+ coarse current and "
                                   "surface flux after the last inner of a group)",
                     "calls_per_step": G * (n_inner + 2), "bytes_per_tally_call": 8 * stride,
                     "ms_per_step": float(cm.item()), "share_of_step": float(cm.item()) / (ms / args.steps)},
            "gpu_launches": int(ln.item()), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": None if bytes_contract is None else bytes_contract / (sweep_ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s",
                         "frac": None if bytes_contract is None else bytes_contract / (sweep_ms * 1e-3) / 1e9 / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "sweep_rchunk_kernel",
                         "launch": "rank 0's sweep-kernel launches of one inner sweep (its angle families: 1/N of the "
                                   "algorithmic bytes)", "ms_per_launch": sweep_ms},
            "cpu_baseline": None,
        }))
    sw.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="pergroup", choices=["pergroup", "batched"])
    ap.add_argument("--boundary", default="gs", choices=["gs", "jacobi"])
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--persistent", type=int, default=0,
                    help="RCHUNK kernel: 1/2 = all plain inners of a sweep call in one cooperative launch (mocb200_options)")
    ap.add_argument("--shard", default="planes", choices=["planes", "angles"],
                    help="N > 1: 'planes' = one N-plane stack, rank r owns plane r (weak scaling, one all-gather per "
                         "sweep(group)); 'angles' = ONE plane, rank r sweeps its angle families, the FSR tally is "
                         "all-reduced after every inner sweep (strong scaling)")
    ap.add_argument("--workload", default="c5g7_2d", choices=sorted(WORKLOADS))
    ap.add_argument("--max-polar", type=int, default=0, help="polar angles bundled per track (0 = library default, 2)")
    ap.add_argument("--cache-groups", type=int, default=0,
                    help="groups the attenuation cache holds at once (0 = all that fit; 1 = what a problem too large for "
                         "the device gets: rebuilt per sweep call)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from mocc_b200 import Sweeper, load_arrays
    from mocc_b200.capi import TALLY_CURRENT

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the MoC sweep has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if rank == 0:
        flat_path = workload_files(args.workload)
    barrier()
    flat_path = workload_files(args.workload)
    arr1 = load_arrays(flat_path)
    S = int(arr1["n_seg_reference"][0])       # per plane
    n_reg1 = int(arr1["n_reg"][0])
    if world > 1 and args.shard == "angles":
        return run_angle_sharded(args, arr1, world, rank, local, barrier)
    arr = stack_planes(arr1, world)            # N > 1: one stack of N planes, rank r owns plane r
    G, n_reg, n_plane = (int(arr[k][0]) for k in ("n_group", "n_reg", "n_plane"))
    bcpg, n_surf, nsp = int(arr["bc_per_group"][0]), int(arr["n_surf"][0]), int(arr["n_surf_plane"][0])
    n_inner = int(arr["n_inner"][0])
    gs = args.boundary == "gs"
    src = synthetic_source(arr, G, n_reg)
    own = [rank] if world > 1 else list(range(n_plane))  # macroplanes of this rank's handle

    sw = Sweeper(arr, device=local, boundary_update=0 if gs else 1, kernel=args.kernel, max_polar=args.max_polar,
                 cache_groups=args.cache_groups, persistent=args.persistent, plane_begin=rank if world > 1 else 0,
                 plane_end=rank + 1 if world > 1 else 0)
    # a dedicated non-default stream: the C ABI treats a NULL stream as "use the handle's own", and
    # torch events only see work on the stream they are recorded on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sw.set_stream(stream.cuda_stream)
    sw.set_xs(0, arr["xs_tr"], xstr_src=arr["xs_tr"], xs_self=arr["xs_self"])
    sw.set_source(0, src)
    sw.set_flux(0, np.ones((G, n_reg)))
    bc0 = np.full((G, bcpg), 1.0 / (4.0 * np.pi))
    for ip in own:
        sw.set_boundary(ip, 0, bc0)

    def step_device():
        if args.mode == "batched":
            sw.sweep(0, G, n_inner=n_inner, tally_mode=TALLY_CURRENT)
        else:
            for g in range(G):
                sw.sweep(g, 1, n_inner=n_inner, tally_mode=TALLY_CURRENT)

    # host buffers of the e2e path live in page-locked memory (the C ABI then copies straight from / into them)
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()
    src = pinned(src)
    flux_h = pinned(np.ones((G, n_reg)))
    bc_h = {ip: pinned(bc0) for ip in own}
    h2d = d2h = 0
    # N > 1: exchange buffers of the per-sweep all-gather (device) and where every rank keeps the whole picture (host)
    n_pack = sw.pack_results_device(0) if world > 1 else 0
    n_pack_max = n_reg1 + 2 * (nsp + int(arr["n_cell_plane"][0]))  # the last plane also carries the top faces
    send = torch.zeros(n_pack_max, dtype=torch.float64, device="cuda") if world > 1 else None
    recv = torch.zeros(world * n_pack_max, dtype=torch.float64, device="cuda") if world > 1 else None
    recv_h = torch.empty(world * n_pack_max, dtype=torch.float64, pin_memory=True) if world > 1 else None
    comm_events = []

    def step_e2e(count=False, time_comm=False):
        nonlocal h2d, d2h
        groups = [range(G)] if args.mode == "batched" else [[g] for g in range(G)]
        blist = [bc_h[ip] if ip in bc_h else None for ip in range(n_plane)]
        for gl in groups:
            # what the C++ plugin does around every sweep(group): one fused upload (source, flux, incoming
            # boundary flux of the handle's planes), the sweep, one fused download
            for g in gl:
                sw.set_sweep_inputs(g, src[g], flux_h[g], [b[g] if b is not None else None for b in blist])
            sw.sweep(gl[0], len(gl), n_inner=n_inner, tally_mode=TALLY_CURRENT)
            for g in gl:
                if world == 1:
                    sw.get_sweep_results(g, flux_h[g], [b[g] if b is not None else None for b in blist], coarse=True)
                    continue
                # N ranks: flux + coarse tallies of every plane to every rank -- packed on the device, ONE NCCL
                # all-gather, one copy to pinned host memory; the rank's own outgoing boundary flux comes back too
                sw.pack_results_device(g, send.data_ptr(), n_pack_max)
                if time_comm:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                dist.all_gather_into_tensor(recv, send)
                if time_comm:
                    e1.record(stream)
                    comm_events.append((e0, e1))
                # the whole picture goes to the host where the (single-process) host solver runs: rank 0; the other
                # ranks keep host mirrors of their own planes only
                if rank == 0:
                    recv_h.copy_(recv, non_blocking=True)
                else:
                    recv_h[:n_pack_max].copy_(send, non_blocking=True)
                sw.get_sweep_results(g, None, [b[g] if b is not None else None for b in blist], coarse=False)
        if count:  # whole job
            h2d = world * G * 8 * (2 * n_reg1 + len(own) * bcpg)
            d2h = G * 8 * (((2 * world - 1) * n_pack_max if world > 1 else n_reg1 + 2 * n_surf) + world * len(own) * bcpg)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = sw.stats()["kernel_launches"]
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for e0, e1 in ev:
        flush.zero_()  # evict the previous step's working set (untimed)
        e0.record(stream)
        step_device()
        e1.record(stream)
    barrier()
    ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    launches = sw.stats()["kernel_launches"] - launches0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    ln = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    updates_step = 2.0 * S * G * n_inner * world  # every rank sweeps its own plane
    value = updates_step * args.steps / (ms * 1e-3)

    # ---- dominant kernel: live CUDA-event time of the sweep kernels of one inner iteration ----
    sw.set_timing(True)
    flush.zero_()
    step_device()
    k_ms, k_n = sw.get_timing()
    sw.set_timing(False)
    sweep_ms = k_ms / max(k_n, 1)                 # one inner sweep of the group set of one call
    groups_per_call = G if args.mode == "batched" else 1
    upd_per_sweep = 2.0 * S * groups_per_call
    peak, peak_src = peaks()
    # contract figure: 6.10 B per update, reference layout, per-group sweep. The resident layout
    # shares geometry between the polar copies; its own compulsory bytes are reported beside it.
    n_useg = int(arr["seg_len"].size)
    n_ray = int(arr1["n_ray_reference"][0])
    n_reg = n_reg1  # the roofline terms below are per plane = per rank
    bytes_contract = BYTES_PER_UPDATE * 2.0 * S if (args.mode == "pergroup" and args.workload == "c5g7_2d") else \
        12.0 * S + groups_per_call * (24.0 * n_reg + 32.0 * n_ray)
    bytes_resident = 12.0 * n_useg + groups_per_call * (24.0 * n_reg + 32.0 * n_ray)
    achieved = bytes_contract / (sweep_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and args.workload == "c5g7_2d":
        traffic = json.load(open(tp)).get(args.mode)

    # ---- end to end through the C ABI with host buffers ----
    step_e2e(count=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = updates_step * args.steps / float(t.item())
    comm = None
    if world > 1:  # device time of the all-gathers of one more step (max over ranks)
        step_e2e(time_comm=True)
        torch.cuda.synchronize()
        cm = torch.tensor([sum(a.elapsed_time(b) for a, b in comm_events)], dtype=torch.float64, device="cuda")
        dist.all_reduce(cm, op=dist.ReduceOp.MAX)
        comm = {"collective": "ncclAllGather of [flux | coarse current | surface flux] slices, one per sweep(group)",
                "calls_per_step": len(comm_events), "bytes_per_call_per_rank": 8 * n_pack_max,
                "ms_per_step": float(cm.item()), "share_of_e2e_step": float(cm.item()) / (float(t.item()) / args.steps * 1e3)}
        assert n_pack <= n_pack_max
    # ---- the same step with the sources built ON THE DEVICE: per step the host sends k (a
    #      call argument), the device computes the fission source and every group's fission + in-scatter source from the
    #      resident flux (Gauss-Seidel over the groups, fixed_source_solver.cpp:102-117); downloads as before ----
    e2e_dev = None
    if world == 1 and all(k in arr for k in ("xs_nf", "xs_ch", "xs_scat")):
        from mocc_b200.capi import material_tables
        fsr_mat, t_nf, t_ch, t_scat = material_tables(arr["xs_nf"], arr["xs_ch"], arr["xs_scat"])
        sw.set_source_xs(fsr_mat, t_nf, t_ch, t_scat)
        blist0 = [bc_h[ip] for ip in range(n_plane)]

        def step_e2e_device_sources():
            sw.fission_source(1.0)
            for g in range(G):
                sw.build_source(g, 1)
                sw.sweep(g, 1, n_inner=n_inner, tally_mode=TALLY_CURRENT)
                sw.get_sweep_results(g, flux_h[g], [b[g] for b in blist0], coarse=True)
        sw.set_flux(0, np.ones((G, n_reg)))
        step_e2e_device_sources()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e_device_sources()
        barrier()
        dt = time.perf_counter() - t0
        e2e_dev = {"value": updates_step * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": 8,
                   "d2h_bytes_per_step": d2h, "materials": int(t_nf.shape[0]),
                   "what": "fission source and group sources built on the device from the resident flux "
                           "(mocb200_fission_source / mocb200_build_source); the host sends k"}
    # the clock sampler covers the device-resident AND the end-to-end timed regions (both under load)
    clocks = sampler.stop() if sampler else None

    # ---- reference CPU sweep on this box's host cores (rank 0, N = 1): one step of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "c5g7_2d" and os.path.exists(REF_TOOL):
        cores = os.cpu_count() or 1
        env = dict(os.environ, OMP_NUM_THREADS=str(cores))
        out = subprocess.run([REF_TOOL, "time", "c5g7_2d.xml", "--cmfd", "--sweeps", "2", "--warmup", "1"],
                             cwd=os.path.join(ROOT, "oracle", "_ref", "inputs"), env=env, capture_output=True,
                             text=True)
        js = [x for x in out.stdout.splitlines() if x.startswith("{")]
        if js:
            r = json.loads(js[-1])
            cpu = {"value": r["updates_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                   "sample": f"two steps (after one warm-up step) of the same workload (7 groups x {r['n_inner']} inners, moc::Current on the "
                             f"last inner): {r['updates']:.3e} updates in {r['seconds']:.2f} s, reference MoCSweeper "
                             f"(OpenMP, unmodified sources in oracle/_ref)"}

    # ---- the plugin inside MOCC's own eigenvalue solve of the same input (rank 0, N = 1): what CudaMoCSweeper::sweep
    #      delivers (pageable Blitz/Eigen arrays, post_sweep on the host) and the time to converge next to the
    #      reference sweeper on this box's host cores ----
    plugin = None
    solve_bin = os.path.join(BIN, "mocc_b200_solve")
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "c5g7_2d" and os.path.exists(solve_bin):
        import re
        import tempfile

        def solve(sweeper_type):
            with tempfile.TemporaryDirectory() as td:
                out = subprocess.run([solve_bin, "c5g7_2d.xml", os.path.join(td, "out.arrays"), "--set",
                                      f"solver/sweeper@type={sweeper_type}"], cwd=os.path.join(BIN, "inputs"),
                                     env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1)),
                                     capture_output=True, text=True)
            m = re.search(r"outers=(\d+) k=([-0-9.e+]+) sweep_seconds=([0-9.e+-]+) device_sweep_ms=([0-9.e+-]+) "
                          r"solve_seconds=([0-9.e+-]+)", out.stdout)
            return None if not m else {"outers": int(m.group(1)), "k": float(m.group(2)), "sweep_s": float(m.group(3)),
                                       "device_sweep_ms": float(m.group(4)), "solve_s": float(m.group(5))}
        sw.synchronize()
        pl, rf = solve("moc_cuda"), solve("moc")
        if pl and rf:
            upd = 2.0 * S * G * n_inner
            plugin = {"e2e_plugin": {"value": upd * pl["outers"] / pl["sweep_s"], "unit": UNIT,
                                     "what": "2 S G n_inner outers / MoC-sweeper timer of mocc_b200_solve (upload, sweeps, "
                                             "download, moc::Current::post_sweep; pageable host arrays)",
                                     "device_only": upd * pl["outers"] / (pl["device_sweep_ms"] * 1e-3)},
                      "time_to_converge_s": {"plugin": pl["solve_s"], "reference": rf["solve_s"],
                                             "reference_threads": os.cpu_count() or 1, "outers_plugin": pl["outers"],
                                             "outers_reference": rf["outers"], "k_plugin": pl["k"], "k_reference": rf["k"],
                                             "dk_pcm": (pl["k"] - rf["k"]) * 1e5, "sweep_s_plugin": pl["sweep_s"],
                                             "sweep_s_reference": rf["sweep_s"]}}

    st = sw.stats()
    kname = {1: "item", 2: "track", 3: "cached", 4: "chunk", 5: "rchunk"}.get(int(st["kernel"]), "?")
    if args.mode == "batched" and kname in ("chunk", "rchunk"):
        kname = "cached"  # group-batched sweeps run the 8-group-lane warp kernel on the same cache
    kfunc = {"item": "sweep_kernel", "track": "sweep_warp_kernel<CACHED=false>", "cached": "sweep_warp_kernel<CACHED=true>",
             "chunk": "sweep_chunk_kernel", "rchunk": "sweep_rchunk_kernel"}[kname]
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64",
            "data": "examples/c5g7_2d.xml geometry and cross sections (flattened on the box); synthetic fixed "
                    "source (fission + in-scatter of a flat unit flux)",
            "config": workload_config(args.workload, args.mode, args.boundary, S, n_reg, G, n_inner, world),
            "arm": {"kernel": kname, "resident_segments": n_useg, "bundled_segments": int(st["swept_segments"]),
                    "max_polar": args.max_polar or 2, "cache_groups": args.cache_groups or G},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "e2e_device_sources": e2e_dev,
            "comm": comm,
            "gpu_launches": int(ln.item()),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": kfunc,
                         "launch": f"the sweep-kernel launches of one inner sweep of {groups_per_call} group(s) "
                                   f"(two boundary phases), averaged over the {n_inner} inners of every sweep call",
                         "ms_per_launch": sweep_ms, "updates_per_launch": upd_per_sweep,
                         "algorithmic_bytes_per_launch": bytes_contract,
                         "resident_layout_bytes_per_launch": bytes_resident,
                         "achieved_resident_layout": bytes_resident / (sweep_ms * 1e-3) / 1e9},
            "cpu_baseline": cpu,
            **(plugin or {}),
        }))
    sw.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
