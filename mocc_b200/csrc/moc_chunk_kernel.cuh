// moc_chunk_kernel.cuh -- the per-group production sweep kernel (sm_100a): ONE SCAN PER TRACK.
//
// Restates the same reference loops as moc_sweep_kernel.cuh
//   sweep1g<CurrentWorker>            src/sweepers/moc/moc_sweeper_kernel.inc.hpp:84-133
//   moc::Current::post_ray            src/sweepers/moc/moc_current_worker.hpp:202-264
//   cmdo::CurrentCorrections::post_ray  src/sweepers/cmdo/correction_worker.hpp:109-205
//   BoundaryCondition::update         src/core/boundary_condition.cpp:155-191
// for one energy group per launch (the reference's sweep(group) contract) on the attenuation
// cache (a = exp_table(-tau) evaluated once per cross-section upload, see exp_cache_kernel).
//
// What a per-group sweep costs on B200 is neither HBM nor FP64 but the SM's load/store data pipe and
// latency: every segment needs one scattered 8-byte q-bar gather and one scattered FP64 reduction
// (profiles/microbench/ubench2_b200.jsonl: 31 resp. 57 cycles per warp instruction that touches 32
// distinct sectors), shared-memory traffic for the lane-owned chunks, and the scan's shuffles. The
// warp-block kernel of moc_sweep_kernel.cuh spends two passes over every 128-segment block, a 5-step
// scan per block and shared-memory transposes: ~275 warp instructions per 32 segments. Here a TEAM
// of NW warps (2 for long tracks) owns one track:
//   * the track's attenuation block (8 P bytes per segment, contiguous in the cache) and its FSR ids are
//     copied into shared memory by TMA bulk copies (cp.async.bulk + mbarrier): read from HBM exactly
//     once per inner sweep, never through registers;
//   * q-bar is gathered with the lanes on CONSECUTIVE segments (neighbouring segments lie in the same
//     pin: few distinct sectors per instruction) by 8-byte cp.async straight into shared memory;
//   * lane i owns the CONTIGUOUS chunk [i L, (i+1) L) of the track (L odd: conflict-free shared-memory
//     strides), composes the affine maps of its chunk serially; ONE butterfly shuffle scan per warp
//     (exclusive prefix: forward, suffix: backward; warp totals exchanged through shared memory) yields
//     the flux entering every chunk in both directions; the lane walks its chunk forward and backward
//     exactly like the reference loop, leaving the summed contribution of every segment in shared memory;
//   * the tally is reduced into global memory again with the lanes on consecutive segments:
//     ONE red.global.add.f64 per segment for both directions and all polar angles.
// Work items are pulled from a global counter; the next item's descriptor, weights and incoming boundary
// flux are prefetched by cp.async into a shared-memory mailbox, its FSR ids by TMA during the current
// item's compute, its attenuations and q-bar behind the current item's reductions.
// Tracks longer than the staging capacity of a team are cut into super-blocks chained by a carried flux
// (a first pass over the super-blocks in reverse order chains the backward flux); the capacity is chosen
// per launch list from the track-length distribution (moc_api.cu: chunk_geometry).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "moc_kernels.cuh"
#include "moc_sweep_kernel.cuh"

namespace mocb200 {

constexpr int kChunkMaxTeam  = 2;  // warps cooperating on one track (3 and 4 measured slower: profiles/r1/tuning.md)
constexpr int kChunkMaxTeams = 16; // teams per CTA
#ifndef MOCB200_CHUNK_ODD_L
#define MOCB200_CHUNK_ODD_L 1
#endif
constexpr int kChunkOddL = MOCB200_CHUNK_ODD_L; // 1: odd chunk length (conflict-free shared-memory strides)
// warps per CTA: the register budget per thread follows (1-2 warps per track: 512 threads, 128 registers;
// 3: 672 threads, 97 registers; 4: 896 threads, 73 registers)
__host__ __device__ constexpr int chunk_max_warps(int nw)
{
    return nw <= 2 ? 16 : (nw == 3 ? 21 : 28);
}

// One (track, polar bundle) of the chunk kernel: everything a warp needs to start the track in ONE
// dependent load (the boundary linkage of BoundaryCondition::update, boundary_condition.cpp:155-191,
// is resolved at set-up). 112 bytes.
struct __align__(16) ChunkUnit {
    int32_t seg_begin; // first segment in the padded segment arrays (multiple of 4)
    int32_t nseg;
    int32_t cpos;      // position of the first (padded) segment inside the list's attenuation cache
    int32_t cross_begin; // crossing lists of the track: forward list, sentinel, backward list, sentinel
    int32_t ang[4];    // sweep-angle indices of the bundle (octants 1-2)
    int32_t in_f[4];   // boundary slot the forward sweep starts from (per polar angle; plane-relative)
    int32_t in_b[4];   // ... the backward sweep
    int32_t out_f[4];  // where the forward exit flux goes: slot >= 0 copy, -(slot+1) write zero (vacuum),
    int32_t out_b[4];  // INT32_MIN leave alone (prescribed); ... backward exit flux
    int32_t n_fw, n_bw; // entries of the forward / backward crossing list (sentinels not counted)
    int32_t pad0, pad1;
};

// bytes of dynamic shared memory one warp needs for `caps` segments: attenuations [caps][P],
// q-bar [caps], summed contributions [caps] (doubles), FSR ids [2][caps] (int32, double-buffered)
__host__ __device__ inline size_t chunk_warp_bytes(int caps, int P, bool tally = false)
{
    // tally variants also stage the track's crossing lists: caps / 2 entries of 8 bytes
    return (size_t)caps * ((size_t)(P + 2) * sizeof(double) + 2 * sizeof(int32_t) + (tally ? 4 : 0));
}

// 8-byte asynchronous copy global -> shared (LDGSTS): the scattered q-bar gather lands in shared
// memory without passing through registers, so a lane keeps its whole stripe in flight at once
__device__ __forceinline__ void cp_async_8(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_16(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <int P> __device__ __forceinline__ void load_ex(const double *exb, int k, double (&e)[P])
{
    if constexpr (P == 2) {
        const double2 v = *reinterpret_cast<const double2 *>(exb + 2 * k);
        e[0] = v.x, e[1] = v.y;
    } else if constexpr (P == 4) {
        const double2 v = *reinterpret_cast<const double2 *>(exb + 4 * k);
        const double2 w = *reinterpret_cast<const double2 *>(exb + 4 * k + 2);
        e[0] = v.x, e[1] = v.y, e[2] = w.x, e[3] = w.y;
    } else {
#pragma unroll
        for (int p = 0; p < P; p++)
            e[p] = exb[k * P + p];
    }
}


// composite affine maps of the lane's chunk: A (both directions), Bf (forward), Bb (backward)
template <int P>
__device__ __forceinline__ void chunk_compose(const double *exb, const double *qb, int lo, int hi, double (&A)[P],
                                              double (&Bf)[P], double (&Bb)[P])
{
#pragma unroll
    for (int p = 0; p < P; p++)
        A[p] = 1.0, Bf[p] = 0.0, Bb[p] = 0.0;
    if (lo >= hi)
        return;
    double ne[P], nq;
    load_ex<P>(exb, lo, ne); // software pipeline: operands of segment k + 1 requested before segment k computes
    nq = qb[lo];
    for (int k = lo; k < hi; k++) {
        double e[P];
#pragma unroll
        for (int p = 0; p < P; p++)
            e[p] = ne[p];
        const double q = nq;
        if (k + 1 < hi) {
            load_ex<P>(exb, k + 1, ne);
            nq = qb[k + 1];
        }
#pragma unroll
        for (int p = 0; p < P; p++) {
            const double bq = q * (1.0 - e[p]);
            Bb[p] = fma(A[p], bq, Bb[p]); // M o m_k: the backward sweep applies the higher segment first
            Bf[p] = fma(e[p], Bf[p], bq); // m_k o M
            A[p] *= e[p];
        }
    }
}

// Scans over the 32 lanes of a warp. In: the map of the lane's own chunk, x -> A x + Bf (forward), A x + Bb
// (backward). Out: the EXCLUSIVE prefix (Ef: the chunks of the lower lanes, forward) and suffix (Eb: the chunks
// of the higher lanes, backward) and, in every lane, the total map of the warp (TA, TF, TB).
// Butterfly form: at step s a lane exchanges the TOTAL map of its 2s-aligned block of s lanes with lane ^ s
// (three values) and extends its prefix (upper half) or suffix (lower half) by the partner block: 6 shuffles
// of 32 bits per polar angle and step instead of the 8 of a Kogge-Stone scan of two pairs, and no final shift
// -- shuffles are a third of the kernel's shared-memory wavefronts.
template <int P>
__device__ __forceinline__ void chunk_scan(int lane, const double (&A)[P], const double (&Bf)[P], const double (&Bb)[P],
                                           double (&EfA)[P], double (&EfB)[P], double (&EbA)[P], double (&EbB)[P],
                                           double (&TA)[P], double (&TF)[P], double (&TB)[P])
{
#pragma unroll
    for (int p = 0; p < P; p++) {
        EfA[p] = 1.0, EfB[p] = 0.0, EbA[p] = 1.0, EbB[p] = 0.0;
        TA[p] = A[p], TF[p] = Bf[p], TB[p] = Bb[p];
    }
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const bool upper = (lane & s) != 0; // the partner block holds the LOWER segments
#pragma unroll
        for (int p = 0; p < P; p++) {
            const double oA = __shfl_xor_sync(0xffffffffu, TA[p], s);
            const double oF = __shfl_xor_sync(0xffffffffu, TF[p], s);
            const double oB = __shfl_xor_sync(0xffffffffu, TB[p], s);
            if (upper) {
                // forward: the lower block comes first, then what I already have below me
                EfB[p] = fma(EfA[p], oF, EfB[p]);
                EfA[p] *= oA;
                // totals of the merged block: forward lower then upper (mine), backward upper (mine) then lower
                TF[p] = fma(TA[p], oF, TF[p]);
                TB[p] = fma(oA, TB[p], oB);
            } else {
                // backward: the upper block comes first, then what I already have above me
                EbB[p] = fma(EbA[p], oB, EbB[p]);
                EbA[p] *= oA;
                TF[p] = fma(oA, TF[p], oF);
                TB[p] = fma(TA[p], oB, TB[p]);
            }
            TA[p] *= oA;
        }
    }
}

// walk the chunk in both directions like the reference loop (kernel:103-129); the two walks are
// independent dependency chains and meet in the middle of the chunk. ab[k] receives the summed tally
// contribution of segment k (both directions, all polar angles).
template <int P>
__device__ __forceinline__ void chunk_walk(const double *exb, const double *qb, double *ab, int lo, int hi,
                                           const double (&wt)[P], double (&psi_f)[P], double (&psi_b)[P])
{
    const int len = hi - lo;
    if (len <= 0)
        return;
    // software pipeline: the shared-memory operands of step j + 1 are requested before step j computes
    double nf[P], nr[P], nqf, nqr;
    load_ex<P>(exb, lo, nf);
    load_ex<P>(exb, hi - 1, nr);
    nqf = qb[lo], nqr = qb[hi - 1];
    for (int j = 0; j < len; j++) {
        const int kf = lo + j, kb = hi - 1 - j; // forward walk at kf, backward walk at kb
        double ef[P], er[P];
#pragma unroll
        for (int p = 0; p < P; p++)
            ef[p] = nf[p], er[p] = nr[p];
        const double qf = nqf, qr = nqr;
        double sf = 0.0, sr = 0.0;
        if (kf > kb) // second visits: add to what the other direction left
            sf = ab[kf], sr = ab[kb];
        if (j + 1 < len) {
            load_ex<P>(exb, kf + 1, nf);
            load_ex<P>(exb, kb - 1, nr);
            nqf = qb[kf + 1], nqr = qb[kb - 1];
        }
#pragma unroll
        for (int p = 0; p < P; p++) {
            const double df = (psi_f[p] - qf) * (1.0 - ef[p]);
            const double dr = (psi_b[p] - qr) * (1.0 - er[p]);
            psi_f[p] -= df;
            psi_b[p] -= dr;
            sf = fma(df, wt[p], sf);
            sr = fma(dr, wt[p], sr);
        }
        if (kf == kb) { // middle segment of an odd chunk: both directions at once
            ab[kf] = sf + sr;
        } else {
            ab[kf] = sf;
            ab[kb] = sr;
        }
    }
}

struct ChunkArgs {
    const ChunkUnit *units;
    int32_t n_units;
    uint32_t *counter;
    const int2 *pinfo; // per plane of the list: {macroplane, its first FSR}
    int32_t n_planes;
    const int32_t *seg_fsr; // padded FSR ids
    const double *wt_v_st;  // [n_plane][n_ang]
    int32_t n_ang, bc_per_group;
    int32_t g_begin, g_count, GP, n_reg;
    const double *q; // group-major [g - g_begin][n_reg]
    double *tally;   // same layout
    const double *bc_in;
    double *bc_out;
    double *scratch; // per team: backward flux entering each super-block of a long track
    int32_t scratch_per_warp;
    const double *cache; // attenuation cache of this list [plane][g][pos][P]
    int64_t list_pseg;
    int32_t cache_groups, cache_g0; // groups per plane in the cache, first group it holds
    int32_t caps; // segments a team stages at once
    int32_t ex_mode; // 0: attenuations by TMA bulk copy, 1: by 16-byte cp.async of all lanes (tuning)
    // coarse-mesh tallies of the last inner (TALLY 1: moc::Current, 2: cmdo::CurrentCorrections)
    const int2 *xptr;   // per 4 padded segments: first forward / backward crossing index
    const Cross *cross; // crossing lists with sentinels
    const double *cur_w, *flx_w; // [n_plane][n_ang][2]
    const int32_t *plane_surf_offset;
    double *current, *surface_flux; // [n_surf][GP]
    double *dsum; // psi_diff per FSR, angle and direction: [g][n_reg][2 n_ang]
    double *ssum; // psi per crossing: [g][plane][n_ang][n_surf_plane][2]
    int32_t n_surf_plane, n_plane_total;
};

// what the tally variants of the walk need beside the staged data
struct ChunkTallyCtx {
    const ChunkArgs *a;
    const int32_t *fb; // FSR ids of the staged block (plane-local)
    const Cross *xl;   // crossing lists of the track (shared-memory copy when it fits, else global)
    int n_fw, n_bw;
    int seg_begin, nseg, k_off; // track position of the staged block
    int plane, first_reg, grel, g;
    int ang[4];
    double cw[4][2], fw[4][2]; // current / surface-flux weights of the bundle's angles (X, Y normal)
    int surf_off;
};

// The walk of the last inner: as chunk_walk, plus moc::Current::post_ray (moc_current_worker.hpp:202-264)
// or cmdo::CurrentCorrections::post_ray (correction_worker.hpp:109-205) at the coarse-surface crossings the
// lane's chunk contains. The two directions are walked one after the other.
template <int P, int TALLY>
__device__ __forceinline__ void chunk_walk_tally(const double *exb, const double *qb, double *ab, int lo, int hi,
                                                 const double (&wt)[P], double (&psi_f)[P], double (&psi_b)[P],
                                                 const ChunkTallyCtx &c)
{
    const ChunkArgs &a = *c.a;
    if (lo >= hi)
        return;
    const int GP       = a.GP;
    const int nslot    = 2 * a.n_ang;
    const int surf_off = c.surf_off;
    auto tally_cross = [&](const Cross &x, const double (&psi)[P], int dir) {
        const int norm = x.surf & 1;
        const int surf = x.surf >> 1;
        const size_t o = (size_t)(surf + surf_off) * GP + c.g;
        double cs = 0.0, fsum = 0.0;
#pragma unroll
        for (int p = 0; p < P; p++) {
            cs   = fma(psi[p], norm ? c.cw[p][1] : c.cw[p][0], cs);
            fsum = fma(psi[p], norm ? c.fw[p][1] : c.fw[p][0], fsum);
        }
        // forward adds, backward subtracts (moc_current_worker.hpp:230-231); the corrections worker also
        // subtracts the backward SURFACE FLUX (correction_worker.hpp:136-137, 194-195)
        atomicAdd(&a.current[o], dir ? -cs : cs);
        atomicAdd(&a.surface_flux[o], (dir && TALLY == 2) ? -fsum : fsum);
        if (TALLY == 2) {
#pragma unroll
            for (int p = 0; p < P; p++) {
                const size_t so = (size_t)c.grel * a.n_plane_total * a.n_ang * a.n_surf_plane * 2 +
                                  (((size_t)c.plane * a.n_ang + c.ang[p]) * a.n_surf_plane + surf) * 2 + dir;
                atomicAdd(&a.ssum[so], psi[p]);
            }
        }
    };
    auto dsum_add = [&](int reg, int p, int dir, double d) {
        const size_t o = (size_t)c.grel * a.n_reg * nslot + (size_t)reg * nslot + c.ang[p] * 2 + dir;
        atomicAdd(&a.dsum[o], d);
    };
    const int nseg = c.nseg;
    // first crossings at or after the chunk's first node, in either walk order (binary search in the lists)
    const int kt_lo = c.k_off + lo, kt_hi = c.k_off + hi; // track positions [kt_lo, kt_hi)
    const Cross *xfl = c.xl, *xbl = c.xl + c.n_fw + 1;
    auto lower_bound = [](const Cross *l, int n, int node) {
        int a0 = 0, a1 = n; // first index with l[i].node >= node (the sentinel at n has node INT32_MAX)
        while (a0 < a1) {
            const int m = (a0 + a1) >> 1;
            if (l[m].node < node)
                a0 = m + 1;
            else
                a1 = m;
        }
        return a0;
    };
    int ci_f = lower_bound(xfl, c.n_fw, kt_lo);
    int ci_b = lower_bound(xbl, c.n_bw, nseg - kt_hi);
    Cross xf = xfl[ci_f], xb = xbl[ci_b];
    auto next_f = [&]() { xf = xfl[++ci_f]; };
    auto next_b = [&]() { xb = xbl[++ci_b]; };
    // ---- both directions interleaved, as in chunk_walk: forward at kf, backward at kb ----
    const int len = hi - lo;
    for (int j = 0; j < len; j++) {
        const int kf = lo + j, kb = hi - 1 - j;
        const int ktf = c.k_off + kf;            // forward flux at the node in front of segment ktf
        const int ktb = c.k_off + kb;
        const int nb  = nseg - 1 - ktb;          // segments walked by the backward sweep so far
        while (xf.node == ktf) {
            tally_cross(xf, psi_f, 0);
            next_f();
        }
        while (xb.node == nb) {
            tally_cross(xb, psi_b, 1);
            next_b();
        }
        double ef[P], er[P];
        load_ex<P>(exb, kf, ef);
        load_ex<P>(exb, kb, er);
        const double qf = qb[kf], qr = qb[kb];
        double sf = 0.0, sr = 0.0;
        if (kf > kb) // second visits: add to what the other direction left
            sf = ab[kf], sr = ab[kb];
#pragma unroll
        for (int p = 0; p < P; p++) {
            const double df = (psi_f[p] - qf) * (1.0 - ef[p]);
            const double dr = (psi_b[p] - qr) * (1.0 - er[p]);
            psi_f[p] -= df;
            psi_b[p] -= dr;
            sf = fma(df, wt[p], sf);
            sr = fma(dr, wt[p], sr);
            if (TALLY == 2) {
                dsum_add(c.fb[kf] + c.first_reg, p, 0, df);
                dsum_add(c.fb[kb] + c.first_reg, p, 1, dr);
            }
        }
        if (kf == kb) {
            ab[kf] = sf + sr;
        } else {
            ab[kf] = sf;
            ab[kb] = sr;
        }
        if (ktf == nseg - 1) { // far end of the ray
            while (xf.node == nseg) {
                tally_cross(xf, psi_f, 0);
                next_f();
            }
        }
        if (ktb == 0) { // near end of the ray
            while (xb.node == nseg) {
                tally_cross(xb, psi_b, 1);
                next_b();
            }
        }
    }
}

// A work item as the team sees it: filled asynchronously (cp.async) one item ahead, in shared memory
struct __align__(16) ChunkWork {
    ChunkUnit u;
    int2 pinfo; // {macroplane, first FSR}
    int32_t ipl, grel; // plane index within the list, group index within the launch
    double wt[4], cf[4], cb[4]; // angle weights, incoming boundary flux (forward, backward)
};

// NW warps ("team") cooperate on one track: 32 NW lanes, each owning one contiguous chunk.
template <int P, int NW, int TALLY>
__global__ void __launch_bounds__(32 * chunk_max_warps(NW), 1) sweep_chunk_kernel(const ChunkArgs a)
{
    constexpr int T = 32 * NW; // lanes of a team
    extern __shared__ __align__(16) double s_dyn[];
    __shared__ uint64_t s_bar[3 * kChunkMaxTeams];
    __shared__ double s_tot[kChunkMaxTeams][kChunkMaxTeam][4][4]; // per team, per warp: Af, Bf, Ab, Bb of the warp's lanes
    __shared__ uint32_t s_w[2 * kChunkMaxTeams];
    __shared__ ChunkWork s_work[kChunkMaxTeams][2];

    const int caps = a.caps;
    const int lane = threadIdx.x & 31;
    const int wid  = threadIdx.x >> 5;
    const int team = wid / NW, wl = wid - team * NW;
    const int tl   = wl * 32 + lane; // lane within the team
    const bool leader = tl == 0;          // work counter, backward exit flux
    const bool loader = tl == T - 32;     // first lane of the team's last warp: issues the TMA copies
    char *wbase    = reinterpret_cast<char *>(s_dyn) + (size_t)team * chunk_warp_bytes(caps, P, TALLY != 0);
    double *exb    = reinterpret_cast<double *>(wbase);
    double *qb     = exb + (size_t)caps * P;
    double *ab     = qb + caps;
    int32_t *fbuf  = reinterpret_cast<int32_t *>(ab + caps); // two FSR-id buffers (plane-local ids)
    Cross *xsm     = reinterpret_cast<Cross *>(fbuf + 2 * caps); // TALLY: the track's crossing lists (caps / 2 entries)
    uint64_t *bar  = &s_bar[3 * team];                       // [0], [1] FSR-id buffers, [2] attenuations
    auto team_sync = [&]() {
        if (NW == 1)
            __syncwarp();
        else
            asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(T) : "memory");
    };
    if (leader) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_init(bar + 2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    team_sync();
    uint32_t par_f = 0u, par_e = 0u; // mbarrier phase parities (bit fi of par_f: FSR-id buffer fi)

    const int GP            = a.GP;
    const uint32_t per_unit = (uint32_t)a.n_planes * (uint32_t)a.g_count;
    const uint32_t total    = (uint32_t)a.n_units * per_unit;
    const int team_global   = blockIdx.x * ((blockDim.x >> 5) / NW) + team;
    double *sc              = a.scratch + (size_t)team_global * a.scratch_per_warp;
    const int32_t *__restrict__ seg_fsr = a.seg_fsr;

    // ---- asynchronous two-level prefetch of a work item into its shared-memory slot ----
    // level 1: descriptor and plane info (addresses depend on the work index only)
    auto prefetch_unit = [&](uint32_t w, ChunkWork *k) {
        if (wl == NW - 1 && lane < 9) {
            const uint32_t unit_id = w / per_unit;
            const uint32_t r       = w - unit_id * per_unit;
            const uint32_t ipl     = r / (uint32_t)a.g_count;
            if (lane < 7)
                cp_async_16(reinterpret_cast<char *>(&k->u) + 16 * lane,
                            reinterpret_cast<const char *>(a.units + unit_id) + 16 * lane);
            else if (lane == 7)
                cp_async_8(&k->pinfo, a.pinfo + ipl);
            else
                k->ipl = (int)ipl, k->grel = (int)(r - ipl * (uint32_t)a.g_count);
        }
    };
    // level 2 (descriptor visible): angle weights and incoming boundary flux. Boundary values read here are
    // never written by the same launch (a launch is one boundary phase / one Jacobi buffer).
    auto prefetch_flux = [&](ChunkWork *k) {
        if (wl == NW - 1 && lane < 12) {
            const int p = lane & 3, kind = lane >> 2;
            if (p < P) {
                const int g     = a.g_begin + k->grel;
                const int plane = k->pinfo.x;
                if (kind == 0)
                    cp_async_8(&k->wt[p], a.wt_v_st + plane * a.n_ang + k->u.ang[p]);
                else {
                    const int slot = kind == 1 ? k->u.in_f[p] : k->u.in_b[p];
                    cp_async_8(kind == 1 ? &k->cf[p] : &k->cb[p],
                               a.bc_in + ((size_t)plane * a.bc_per_group + slot) * GP + g);
                }
            }
        }
    };
    // TMA bulk copies of one (super-)block: FSR ids into buffer fi, attenuations into exb
    auto issue_fsr = [&](int fi, const ChunkWork *k, int k_off, int n) {
        if (loader) {
            const uint32_t bytes = (uint32_t)((n + 3) & ~3) * 4u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar + fi, bytes);
            bulk_g2s(fbuf + fi * caps, seg_fsr + k->u.seg_begin + k_off, bytes, bar + fi);
        }
    };
    auto issue_ex = [&](const ChunkWork *k, int k_off, int n) {
        if (a.ex_mode == 1) {
            const int g        = a.g_begin + k->grel;
            const double *ex_g = a.cache + (((size_t)k->ipl * a.cache_groups + (g - a.cache_g0)) * a.list_pseg + k->u.cpos + k_off) * P;
            const int n16      = ((n + 3) & ~3) * P / 2;
#pragma unroll 4
            for (int i = tl; i < n16; i += T)
                cp_async_16(exb + 2 * i, ex_g + 2 * i);
        } else if (loader) {
            const int g        = a.g_begin + k->grel;
            const double *ex_g = a.cache + (((size_t)k->ipl * a.cache_groups + (g - a.cache_g0)) * a.list_pseg + k->u.cpos) * P;
            const uint32_t bytes = (uint32_t)((n + 3) & ~3) * (uint32_t)P * 8u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar + 2, bytes);
            const char *src = reinterpret_cast<const char *>(ex_g + (size_t)k_off * P);
            for (uint32_t off = 0; off < bytes; off += 16384u) // <= 16 KB per bulk copy
                bulk_g2s(reinterpret_cast<char *>(exb) + off, src + off, min(16384u, bytes - off), bar + 2);
        }
    };
    // asynchronous striped q-bar gather (lanes on consecutive segments) straight into shared memory
    auto gather_q = [&](int fi, const ChunkWork *k, int n) {
        mbar_wait(bar + fi, (par_f >> fi) & 1u);
        par_f ^= 1u << fi;
        const double *qf  = a.q + (size_t)k->grel * a.n_reg + k->pinfo.y;
        const int32_t *fb = fbuf + fi * caps;
#pragma unroll 8
        for (int i = tl; i < n; i += T)
            cp_async_8(qb + i, qf + fb[i]);
        if (TALLY != 0) { // crossing lists of the track (both sentinels included) when they fit
            const int nx = k->u.n_fw + k->u.n_bw + 2;
            if (nx <= caps / 2) {
                const Cross *src = a.cross + k->u.cross_begin;
                for (int i = tl; i < nx; i += T)
                    cp_async_8(xsm + i, src + i);
            }
        }
    };
    auto wait_staged = [&]() {
        if (a.ex_mode == 0) {
            mbar_wait(bar + 2, par_e);
            par_e ^= 1u;
        }
        cp_async_wait_all();
        team_sync();
    };
    auto reduce_tally = [&](int fi, const ChunkWork *k, int n) {
        double *tf        = a.tally + (size_t)k->grel * a.n_reg + k->pinfo.y;
        const int32_t *fb = fbuf + fi * caps;
#pragma unroll 8
        for (int i = tl; i < n; i += T)
            atomicAdd(&tf[fb[i]], ab[i]);
    };
    // One staged (super-)block. cf: forward flux entering the block (team-uniform), eb: backward flux entering
    // it from the far side. Leaves the contributions in ab; returns the forward flux leaving the block
    // (valid in the last lane of the team) and the backward flux leaving it (valid in team lane 0).
    auto block = [&](int n, int k_off, int fi, const ChunkWork *k, const double (&wt)[P], const double (&cf)[P],
                     const double (&eb)[P], double (&out_fwd)[P], double (&out_bwd)[P]) {
        const int L  = ((n + T - 1) / T) | kChunkOddL;
        const int lo = min(tl * L, n), hi = min(lo + L, n);
        ChunkTallyCtx c;
        if (TALLY != 0) { // requested first: the weights' global latency hides behind compose and scan
            c.a = &a, c.fb = fbuf + fi * caps;
            c.n_fw = k->u.n_fw, c.n_bw = k->u.n_bw;
            c.xl = (c.n_fw + c.n_bw + 2 <= caps / 2) ? xsm : a.cross + k->u.cross_begin;
            c.seg_begin = k->u.seg_begin, c.nseg = k->u.nseg, c.k_off = k_off;
            c.plane = k->pinfo.x, c.first_reg = k->pinfo.y, c.grel = k->grel, c.g = a.g_begin + k->grel;
#pragma unroll
            for (int p = 0; p < 4; p++)
                c.ang[p] = k->u.ang[p];
#pragma unroll
            for (int p = 0; p < P; p++) {
                const size_t o = ((size_t)c.plane * a.n_ang + c.ang[p]) * 2;
                c.cw[p][0] = a.cur_w[o], c.cw[p][1] = a.cur_w[o + 1];
                c.fw[p][0] = a.flx_w[o], c.fw[p][1] = a.flx_w[o + 1];
            }
            c.surf_off = a.plane_surf_offset[c.plane];
        }
        double A[P], Bf[P], Bb[P], EfA[P], EfB[P], EbA[P], EbB[P], TA[P], TF[P], TB[P];
        chunk_compose<P>(exb, qb, lo, hi, A, Bf, Bb);
        chunk_scan<P>(lane, A, Bf, Bb, EfA, EfB, EbA, EbB, TA, TF, TB);
        double cfw[P], ebw[P];
#pragma unroll
        for (int p = 0; p < P; p++)
            cfw[p] = cf[p], ebw[p] = eb[p];
        if (NW > 1) { // maps of the other warps of the team: forward through the lower, backward through the higher ones
            if (lane == 0) { // every lane holds the warp's total map
#pragma unroll
                for (int p = 0; p < P; p++) {
                    s_tot[team][wl][p][0] = TA[p], s_tot[team][wl][p][1] = TF[p];
                    s_tot[team][wl][p][2] = TA[p], s_tot[team][wl][p][3] = TB[p];
                }
            }
            team_sync();
#pragma unroll
            for (int w = 0; w < NW - 1; w++) {
                if (w < wl) {
#pragma unroll
                    for (int p = 0; p < P; p++)
                        cfw[p] = fma(s_tot[team][w][p][0], cfw[p], s_tot[team][w][p][1]);
                }
            }
#pragma unroll
            for (int w = NW - 1; w > 0; w--) {
                if (w > wl) {
#pragma unroll
                    for (int p = 0; p < P; p++)
                        ebw[p] = fma(s_tot[team][w][p][2], ebw[p], s_tot[team][w][p][3]);
                }
            }
        }
        double psi_f[P], psi_b[P];
#pragma unroll
        for (int p = 0; p < P; p++) {
            psi_f[p]   = fma(EfA[p], cfw[p], EfB[p]); // flux entering this lane's chunk, forward
            psi_b[p]   = fma(EbA[p], ebw[p], EbB[p]); // ... backward
            out_fwd[p] = fma(TA[p], cfw[p], TF[p]);   // flux leaving the warp's lanes, forward (every lane)
        }
        if (TALLY == 0) {
            chunk_walk<P>(exb, qb, ab, lo, hi, wt, psi_f, psi_b);
        } else {
            chunk_walk_tally<P, TALLY>(exb, qb, ab, lo, hi, wt, psi_f, psi_b, c);
        }
#pragma unroll
        for (int p = 0; p < P; p++)
            out_bwd[p] = psi_b[p];
    };

    // ---- software pipeline over the work items: counter two ahead, descriptor and boundary flux one ahead ----
    // plain PTX atomic: the compiler's warp-aggregated atomicAdd would broadcast (and so wait for) the result at once
    auto fetch = [&]() -> uint32_t {
        uint32_t w = 0u;
        if (leader)
            asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(w) : "l"(a.counter) : "memory");
        return w;
    };
    uint32_t share_slot = 0u;
    auto share = [&](uint32_t w) -> uint32_t { // leader's value to the whole team (one barrier; alternating slots)
        if (NW == 1)
            return __shfl_sync(0xffffffffu, w, 0);
        if (leader)
            s_w[2 * team + share_slot] = w;
        team_sync();
        const uint32_t r = s_w[2 * team + share_slot];
        share_slot ^= 1u;
        return r;
    };
    uint32_t w_cur = share(fetch());
    uint32_t w_nxt = share(fetch());
    int cs = 0; // slot of `cur` in s_work[team]
    if (w_cur < total) {
        prefetch_unit(w_cur, &s_work[team][cs]);
        cp_async_wait_all();
        team_sync();
        prefetch_flux(&s_work[team][cs]);
        cp_async_wait_all();
        team_sync();
    }
    bool staged = false; // attenuations + q-bar of `cur` already on their way (issued by the previous item)
    int fi      = 0;     // FSR-id buffer of `cur`

    while (w_cur < total) {
        uint32_t w_nn_raw; // work index two items ahead: fetched once the staging loads have drained, consumed at the end
        const ChunkWork *cur = &s_work[team][cs];
        ChunkWork *nxt       = &s_work[team][cs ^ 1];
        const bool have_nxt  = w_nxt < total;
        if (have_nxt)
            prefetch_unit(w_nxt, nxt);

        const int nseg = cur->u.nseg;
        double cf_out[P], cb_out[P];

        if (nseg <= caps) {
            // ================= the whole track fits: one staged block, next item prefetched =================
            if (!staged) {
                issue_fsr(fi, cur, 0, nseg);
                issue_ex(cur, 0, nseg);
                gather_q(fi, cur, nseg);
            }
            wait_staged(); // also: the next item's descriptor and this item's boundary flux have landed
            w_nn_raw = fetch();
            const bool pre = have_nxt && nxt->u.nseg <= caps;
            if (pre)
                issue_fsr(fi ^ 1, nxt, 0, nxt->u.nseg);
            if (have_nxt)
                prefetch_flux(nxt);
            double wt[P], cf[P], cb[P];
#pragma unroll
            for (int p = 0; p < P; p++)
                wt[p] = cur->wt[p], cf[p] = cur->cf[p], cb[p] = cur->cb[p];
            block(nseg, 0, fi, cur, wt, cf, cb, cf_out, cb_out);
            team_sync();
            if (pre) { // exb and qb are free again: stage the next track behind this one's reductions
                issue_ex(nxt, 0, nxt->u.nseg);
                gather_q(fi ^ 1, nxt, nxt->u.nseg);
            }
            reduce_tally(fi, cur, nseg);
            staged = pre;
            if (pre)
                fi ^= 1;
        } else {
            // ================= long track: super-blocks of caps segments chained by a carried flux =================
            cp_async_wait_all();
            team_sync();
            w_nn_raw = fetch();
            if (have_nxt)
                prefetch_flux(nxt);
            const int nsb = (nseg + caps - 1) / caps;
            double wt[P], cb[P], cf[P];
#pragma unroll
            for (int p = 0; p < P; p++)
                wt[p] = cur->wt[p], cf[p] = cur->cf[p], cb[p] = cur->cb[p];
            for (int sb = nsb - 1; sb >= 1; --sb) { // pass A: backward flux entering each super-block
                const int n = min(caps, nseg - sb * caps);
                issue_fsr(fi, cur, sb * caps, n);
                issue_ex(cur, sb * caps, n);
                gather_q(fi, cur, n);
                wait_staged();
                const int L  = ((n + T - 1) / T) | kChunkOddL;
                const int lo = min(tl * L, n), hi = min(lo + L, n);
                if (leader) {
#pragma unroll
                    for (int p = 0; p < P; p++)
                        sc[sb * P + p] = cb[p];
                }
                double A[P], Bf[P], B[P];
                chunk_compose<P>(exb, qb, lo, hi, A, Bf, B);
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) { // ordered butterfly: total = L_0 o L_1 o ... o L_31
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        const double Ao = __shfl_xor_sync(0xffffffffu, A[p], s);
                        const double Bo = __shfl_xor_sync(0xffffffffu, B[p], s);
                        if (lane & s) // partner holds the lower segments: partner o mine
                            B[p] = fma(Ao, B[p], Bo);
                        else // mine o partner
                            B[p] = fma(A[p], Bo, B[p]);
                        A[p] *= Ao;
                    }
                }
                if (NW > 1) {
                    if (lane == 0) {
#pragma unroll
                        for (int p = 0; p < P; p++)
                            s_tot[team][wl][p][2] = A[p], s_tot[team][wl][p][3] = B[p];
                    }
                    team_sync();
#pragma unroll
                    for (int w = NW - 1; w >= 0; w--) {
#pragma unroll
                        for (int p = 0; p < P; p++)
                            cb[p] = fma(s_tot[team][w][p][2], cb[p], s_tot[team][w][p][3]);
                    }
                } else {
#pragma unroll
                    for (int p = 0; p < P; p++)
                        cb[p] = fma(A[p], cb[p], B[p]);
                }
                team_sync();
            }
            for (int sb = 0; sb < nsb; ++sb) { // pass B: forward chain, both walks, tally
                const int n = min(caps, nseg - sb * caps);
                issue_fsr(fi, cur, sb * caps, n);
                issue_ex(cur, sb * caps, n);
                gather_q(fi, cur, n);
                wait_staged();
                double eb[P], of[P], ob[P];
#pragma unroll
                for (int p = 0; p < P; p++)
                    eb[p] = sb > 0 ? sc[sb * P + p] : cb[p];
                block(n, sb * caps, fi, cur, wt, cf, eb, of, ob);
#pragma unroll
                for (int p = 0; p < P; p++) {
                    if (NW > 1) { // the last lane of the team holds the flux leaving the block
                        if (tl == T - 1)
                            s_tot[team][0][p][0] = of[p];
                    } else {
                        cf[p] = __shfl_sync(0xffffffffu, of[p], 31);
                    }
                    if (sb == 0)
                        cb_out[p] = ob[p];
                }
                team_sync();
                if (NW > 1) {
#pragma unroll
                    for (int p = 0; p < P; p++)
                        cf[p] = s_tot[team][0][p][0];
                }
                reduce_tally(fi, cur, n);
                team_sync();
            }
#pragma unroll
            for (int p = 0; p < P; p++)
                cf_out[p] = cf[p];
            staged = false;
        }

        // ---- outgoing boundary flux, written where BoundaryCondition::update would copy it: the forward
        //      exit flux by the last lane of the team, the backward exit flux by its first lane ----
        if (tl == T - 1 || leader) {
            const int g       = a.g_begin + cur->grel;
            double *bc_out_pl = a.bc_out + (size_t)cur->pinfo.x * a.bc_per_group * GP + g;
            if (tl == T - 1) {
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const int o = cur->u.out_f[p];
                    if (o != INT32_MIN)
                        bc_out_pl[(size_t)(o >= 0 ? o : -(o + 1)) * GP] = o >= 0 ? cf_out[p] : 0.0;
                }
            }
            if (leader) {
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const int o = cur->u.out_b[p];
                    if (o != INT32_MIN)
                        bc_out_pl[(size_t)(o >= 0 ? o : -(o + 1)) * GP] = o >= 0 ? cb_out[p] : 0.0;
                }
            }
        }
        const uint32_t w_nn = share(w_nn_raw); // also fences the team before buffers and slots are reused
        w_cur = w_nxt, w_nxt = w_nn;
        cs ^= 1;
    }
}

} // namespace mocb200
