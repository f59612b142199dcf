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
// What bounds a per-group sweep on B200 is neither HBM nor FP64 but the SM's load/store data
// pipe: every segment needs one scattered 8-byte q-bar gather and one scattered FP64 reduction
// (profiles/microbench/ubench2_b200.jsonl: 31 resp. 57 cycles per warp instruction that touches 32
// distinct sectors). The warp-block kernel of moc_sweep_kernel.cuh adds to that two passes over
// every 128-segment block, a 5-step shuffle scan per block and shared-memory transposes: ~275
// warp instructions per 32 segments. Here a warp still owns one track, but
//   * the track's attenuation stream (8 P bytes per segment, contiguous in the cache) is copied
//     into shared memory by ONE TMA bulk copy (cp.async.bulk + mbarrier): it never touches the
//     load/store pipe on the way in and is read from HBM exactly once per inner sweep;
//   * q-bar is gathered with the lanes on 32 CONSECUTIVE segments (neighbouring segments lie in
//     the same pin: few distinct sectors per instruction) and parked in shared memory;
//   * lane i then owns the CONTIGUOUS chunk [i L, (i+1) L) of the track (L odd: conflict-free
//     shared-memory strides), composes the affine maps of its chunk serially, ONE shuffle scan
//     per track (prefix: forward, suffix: backward) yields the flux entering every chunk in both
//     directions, and the lane walks its chunk forward and backward exactly like the reference
//     loop, leaving the summed contribution of every segment in shared memory;
//   * the tally is reduced into global memory again with the lanes on consecutive segments:
//     ONE red.global.add.f64 per segment for both directions and all polar angles.
// Tracks longer than the shared-memory capacity of a warp are cut into super-blocks chained by
// a carried flux (a first pass over the super-blocks in reverse order chains the backward flux).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "moc_kernels.cuh"
#include "moc_sweep_kernel.cuh"

namespace mocb200 {

constexpr int kChunkMaxWarps = 12; // 384 threads: up to 170 registers per thread (occupancy is shared-memory bound)

// One (track, polar bundle) of the chunk kernel: everything a warp needs to start the track in ONE
// dependent load (the boundary linkage of BoundaryCondition::update, boundary_condition.cpp:155-191,
// is resolved at set-up). 96 bytes.
struct __align__(16) ChunkUnit {
    int32_t seg_begin; // first segment in the padded segment arrays (multiple of 4)
    int32_t nseg;
    int32_t cpos;      // position of the first (padded) segment inside the list's attenuation cache
    int32_t pad;
    int32_t ang[4];    // sweep-angle indices of the bundle (octants 1-2)
    int32_t in_f[4];   // boundary slot the forward sweep starts from (per polar angle; plane-relative)
    int32_t in_b[4];   // ... the backward sweep
    int32_t out_f[4];  // where the forward exit flux goes: slot >= 0 copy, -(slot+1) write zero (vacuum),
    int32_t out_b[4];  // INT32_MIN leave alone (prescribed); ... backward exit flux
};

// bytes of dynamic shared memory one warp needs for `caps` segments: attenuations [caps][P],
// q-bar [caps], summed contributions [caps] (doubles), FSR ids [2][caps] (int32, double-buffered)
__host__ __device__ inline size_t chunk_warp_bytes(int caps, int P)
{
    return (size_t)caps * ((size_t)(P + 2) * sizeof(double) + 2 * sizeof(int32_t));
}

// 8-byte asynchronous copy global -> shared (LDGSTS): the scattered q-bar gather lands in shared
// memory without passing through registers, so a lane keeps its whole stripe in flight at once
__device__ __forceinline__ void cp_async_8(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
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


// composite affine map of the lane's chunk for the backward direction only (pass A of long tracks)
template <int P>
__device__ __forceinline__ void chunk_compose_bwd(const double *exb, const double *qb, int lo, int hi, double (&A)[P],
                                                  double (&B)[P])
{
#pragma unroll
    for (int p = 0; p < P; p++)
        A[p] = 1.0, B[p] = 0.0;
    for (int k = lo; k < hi; k++) {
        double e[P];
        load_ex<P>(exb, k, e);
        const double q = qb[k];
#pragma unroll
        for (int p = 0; p < P; p++) {
            const double bq = q * (1.0 - e[p]);
            B[p] = fma(A[p], bq, B[p]); // M o m_k: the backward sweep applies the higher segment first
            A[p] *= e[p];
        }
    }
}

// One staged (super-)block: compose the lane chunks, scan, walk both directions. On entry cf is the
// forward flux entering the block and eb the backward flux entering it (from the far side); on exit cf
// is the forward flux leaving the block, psi_b (lane 0) the backward flux leaving it, and ab[k] holds
// the summed tally contribution of segment k.
template <int P>
__device__ __forceinline__ void chunk_block(const double *exb, const double *qb, double *ab, int lane, int lo, int hi,
                                            const double (&wt)[P], double (&cf)[P], const double (&eb)[P],
                                            double (&psi_b)[P])
{
    double A[P], Bf[P], Bb[P];
#pragma unroll
    for (int p = 0; p < P; p++)
        A[p] = 1.0, Bf[p] = 0.0, Bb[p] = 0.0;
#pragma unroll 2
    for (int k = lo; k < hi; k++) {
        double e[P];
        load_ex<P>(exb, k, e);
        const double q = qb[k];
#pragma unroll
        for (int p = 0; p < P; p++) {
            const double bq = q * (1.0 - e[p]);
            Bb[p] = fma(A[p], bq, Bb[p]); // M o m_k
            Bf[p] = fma(e[p], Bf[p], bq); // m_k o M
            A[p] *= e[p];
        }
    }
    // ---- one inclusive scan over the lanes: prefix (forward), suffix (backward) ----
    double psi_f[P];
    {
        double Af[P], Ab[P];
#pragma unroll
        for (int p = 0; p < P; p++)
            Af[p] = A[p], Ab[p] = A[p];
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
#pragma unroll
            for (int p = 0; p < P; p++) {
                const double Ae = __shfl_up_sync(0xffffffffu, Af[p], s);
                const double Be = __shfl_up_sync(0xffffffffu, Bf[p], s);
                const double Ah = __shfl_down_sync(0xffffffffu, Ab[p], s);
                const double Bh = __shfl_down_sync(0xffffffffu, Bb[p], s);
                if (lane >= s) { // mine o earlier
                    Bf[p] = fma(Af[p], Be, Bf[p]);
                    Af[p] *= Ae;
                }
                if (lane + s < 32) { // mine o higher
                    Bb[p] = fma(Ab[p], Bh, Bb[p]);
                    Ab[p] *= Ah;
                }
            }
        }
#pragma unroll
        for (int p = 0; p < P; p++) {
            const double out_f = fma(Af[p], cf[p], Bf[p]); // flux leaving this lane's chunk, forward
            const double out_b = fma(Ab[p], eb[p], Bb[p]); // ... backward
            const double in_f  = __shfl_up_sync(0xffffffffu, out_f, 1);
            const double in_b  = __shfl_down_sync(0xffffffffu, out_b, 1);
            psi_f[p] = lane == 0 ? cf[p] : in_f;
            psi_b[p] = lane == 31 ? eb[p] : in_b;
            cf[p]    = __shfl_sync(0xffffffffu, out_f, 31); // carried to the next super-block
        }
    }
    // ---- walk the chunk in both directions like the reference loop (kernel:103-129); the two walks
    //      are independent dependency chains and meet in the middle of the chunk ----
    int kf = lo, kb = hi - 1;
    for (; kf < kb; ++kf, --kb) { // first visit of both segments
        double ef[P], er[P];
        load_ex<P>(exb, kf, ef);
        load_ex<P>(exb, kb, er);
        const double qf = qb[kf], qr = qb[kb];
        double sf = 0.0, sr = 0.0;
#pragma unroll
        for (int p = 0; p < P; p++) {
            const double df = (psi_f[p] - qf) * (1.0 - ef[p]);
            const double dr = (psi_b[p] - qr) * (1.0 - er[p]);
            psi_f[p] -= df;
            psi_b[p] -= dr;
            sf = fma(df, wt[p], sf);
            sr = fma(dr, wt[p], sr);
        }
        ab[kf] = sf;
        ab[kb] = sr;
    }
    if (kf == kb) { // middle segment of an odd chunk: both directions at once
        double em[P];
        load_ex<P>(exb, kf, em);
        const double qm = qb[kf];
        double s = 0.0;
#pragma unroll
        for (int p = 0; p < P; p++) {
            const double df = (psi_f[p] - qm) * (1.0 - em[p]);
            const double dr = (psi_b[p] - qm) * (1.0 - em[p]);
            psi_f[p] -= df;
            psi_b[p] -= dr;
            s = fma(df, wt[p], s);
            s = fma(dr, wt[p], s);
        }
        ab[kf] = s;
        ++kf, --kb;
    }
    for (; kf < hi; ++kf, --kb) { // second visits: add to what the other direction left
        double ef[P], er[P];
        load_ex<P>(exb, kf, ef);
        load_ex<P>(exb, kb, er);
        const double qf = qb[kf], qr = qb[kb];
        double sf = ab[kf], sr = ab[kb];
#pragma unroll
        for (int p = 0; p < P; p++) {
            const double df = (psi_f[p] - qf) * (1.0 - ef[p]);
            const double dr = (psi_b[p] - qr) * (1.0 - er[p]);
            psi_f[p] -= df;
            psi_b[p] -= dr;
            sf = fma(df, wt[p], sf);
            sr = fma(dr, wt[p], sr);
        }
        ab[kf] = sf;
        ab[kb] = sr;
    }
}

struct ChunkArgs {
    const ChunkUnit *units;
    int32_t n_units;
    uint32_t *counter;
    const int32_t *planes;
    int32_t n_planes;
    const int32_t *seg_fsr; // padded FSR ids
    const double *wt_v_st;  // [n_plane][n_ang]
    const int32_t *plane_first_reg;
    int32_t n_ang, bc_per_group;
    int32_t g_begin, g_count, GP, n_reg;
    const double *q; // group-major [g - g_begin][n_reg]
    double *tally;   // same layout
    const double *bc_in;
    double *bc_out;
    double *scratch; // per warp: backward flux entering each super-block of a long track
    int32_t scratch_per_warp;
    const double *cache; // attenuation cache of this list [plane][g][pos][P]
    int64_t list_pseg;
    int32_t cache_groups;
    int32_t caps; // segments a warp stages at once (32 x odd)
};

template <int P>
__global__ void __launch_bounds__(32 * kChunkMaxWarps, 1) sweep_chunk_kernel(const ChunkArgs a)
{
    extern __shared__ __align__(16) double s_dyn[];
    __shared__ uint64_t s_bar[3 * kChunkMaxWarps];

    const int caps = a.caps;
    const int lane = threadIdx.x & 31;
    const int wid  = threadIdx.x >> 5;
    char *wbase    = reinterpret_cast<char *>(s_dyn) + (size_t)wid * chunk_warp_bytes(caps, P);
    double *exb    = reinterpret_cast<double *>(wbase);
    double *qb     = exb + (size_t)caps * P;
    double *ab     = qb + caps;
    int32_t *fbuf  = reinterpret_cast<int32_t *>(ab + caps); // two FSR-id buffers (plane-local ids)
    uint64_t *bar  = &s_bar[3 * wid];                        // [0], [1] FSR-id buffers, [2] attenuations
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_init(bar + 2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t par_f = 0u, par_e = 0u; // mbarrier phase parities (bit fi of par_f: FSR-id buffer fi)

    const int GP            = a.GP;
    const uint32_t per_unit = (uint32_t)a.n_planes * (uint32_t)a.g_count;
    const uint32_t total    = (uint32_t)a.n_units * per_unit;
    const int warp_global   = blockIdx.x * (blockDim.x >> 5) + wid;
    double *sc              = a.scratch + (size_t)warp_global * a.scratch_per_warp;
    const int32_t *__restrict__ seg_fsr = a.seg_fsr;

    auto next_work = [&]() -> uint32_t {
        uint32_t w = 0;
        if (lane == 0)
            w = atomicAdd(a.counter, 1u);
        return __shfl_sync(0xffffffffu, w, 0);
    };
    // the part of a work item that is prefetched one item ahead
    struct Work {
        int4 d0, d1, d2, d3, d4, d5; // ChunkUnit
        int plane, first_reg, ipl, grel;
    };
    auto load_work = [&](uint32_t w, Work &k) {
        const int unit_id = (int)(w / per_unit);
        const uint32_t r  = w - (uint32_t)unit_id * per_unit;
        k.ipl             = (int)(r / (uint32_t)a.g_count);
        k.grel            = (int)(r - (uint32_t)k.ipl * (uint32_t)a.g_count);
        const int4 *u     = reinterpret_cast<const int4 *>(a.units + unit_id);
        k.d0 = u[0], k.d1 = u[1], k.d2 = u[2], k.d3 = u[3], k.d4 = u[4], k.d5 = u[5];
        k.plane     = a.planes[k.ipl];
        k.first_reg = a.plane_first_reg[k.plane];
    };
    auto ex_of = [&](const Work &k) -> const double * {
        return a.cache + (((size_t)k.ipl * a.cache_groups + (a.g_begin + k.grel)) * a.list_pseg + k.d0.z) * P;
    };
    // TMA bulk copies of one (super-)block: FSR ids into buffer fi, attenuations into exb
    auto issue_fsr = [&](int fi, const Work &k, int k_off, int n) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)((n + 3) & ~3) * 4u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar + fi, bytes);
            bulk_g2s(fbuf + fi * caps, seg_fsr + k.d0.x + k_off, bytes, bar + fi);
        }
    };
    auto issue_ex = [&](const Work &k, int k_off, int n) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)((n + 3) & ~3) * (uint32_t)P * 8u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar + 2, bytes);
            const char *src = reinterpret_cast<const char *>(ex_of(k) + (size_t)k_off * P);
            for (uint32_t off = 0; off < bytes; off += 16384u) // <= 16 KB per bulk copy
                bulk_g2s(reinterpret_cast<char *>(exb) + off, src + off, min(16384u, bytes - off), bar + 2);
        }
    };
    // asynchronous striped q-bar gather (lanes on consecutive segments) straight into shared memory
    auto gather_q = [&](int fi, const Work &k, int n) {
        mbar_wait(bar + fi, (par_f >> fi) & 1u);
        par_f ^= 1u << fi;
        const double *qf  = a.q + (size_t)k.grel * a.n_reg + k.first_reg;
        const int32_t *fb = fbuf + fi * caps;
#pragma unroll 4
        for (int i = lane; i < n; i += 32)
            cp_async_8(qb + i, qf + fb[i]);
    };
    auto wait_staged = [&]() {
        mbar_wait(bar + 2, par_e);
        par_e ^= 1u;
        cp_async_wait_all();
        __syncwarp();
    };
    auto reduce_tally = [&](int fi, const Work &k, int n) {
        double *tf        = a.tally + (size_t)k.grel * a.n_reg + k.first_reg;
        const int32_t *fb = fbuf + fi * caps;
#pragma unroll 4
        for (int i = lane; i < n; i += 32)
            atomicAdd(&tf[fb[i]], ab[i]);
    };

    // ---- software pipeline over the work items: counter two ahead, descriptor one ahead ----
    uint32_t w_cur = next_work();
    uint32_t w_nxt = next_work();
    Work cur, nxt;
    if (w_cur < total)
        load_work(w_cur, cur);
    bool staged = false; // attenuations + q-bar of `cur` already on their way (issued by the previous item)
    int fi      = 0;     // FSR-id buffer of `cur`

    while (w_cur < total) {
        const uint32_t w_nn = next_work();
        if (w_nxt < total)
            load_work(w_nxt, nxt);

        const int nseg = cur.d0.y;
        const int g    = a.g_begin + cur.grel;
        const int ang[4]   = {cur.d1.x, cur.d1.y, cur.d1.z, cur.d1.w};
        const int in_f[4]  = {cur.d2.x, cur.d2.y, cur.d2.z, cur.d2.w};
        const int in_b[4]  = {cur.d3.x, cur.d3.y, cur.d3.z, cur.d3.w};
        const int out_f[4] = {cur.d4.x, cur.d4.y, cur.d4.z, cur.d4.w};
        const int out_b[4] = {cur.d5.x, cur.d5.y, cur.d5.z, cur.d5.w};
        double wt[P], cf[P], cb[P];
        const double *bc_in_pl = a.bc_in + (size_t)cur.plane * a.bc_per_group * GP;
#pragma unroll
        for (int p = 0; p < P; p++) {
            wt[p] = a.wt_v_st[cur.plane * a.n_ang + ang[p]];
            cf[p] = bc_in_pl[(size_t)in_f[p] * GP + g];
            cb[p] = bc_in_pl[(size_t)in_b[p] * GP + g];
        }

        if (nseg <= caps) {
            // ================= the whole track fits: one staged block, next item prefetched =================
            if (!staged) {
                issue_fsr(fi, cur, 0, nseg);
                issue_ex(cur, 0, nseg);
                gather_q(fi, cur, nseg);
            }
            const bool pre = w_nxt < total && nxt.d0.y <= caps;
            if (pre)
                issue_fsr(fi ^ 1, nxt, 0, nxt.d0.y);
            wait_staged();
            const int L  = ((nseg + 31) >> 5) | 1;
            const int lo = min(lane * L, nseg), hi = min(lo + L, nseg);
            double psi_b[P];
            chunk_block<P>(exb, qb, ab, lane, lo, hi, wt, cf, cb, psi_b);
#pragma unroll
            for (int p = 0; p < P; p++)
                cb[p] = psi_b[p]; // lane 0: backward flux leaving the ray
            __syncwarp();
            if (pre) { // exb and qb are free again: stage the next track behind this one's reductions
                issue_ex(nxt, 0, nxt.d0.y);
                gather_q(fi ^ 1, nxt, nxt.d0.y);
            }
            reduce_tally(fi, cur, nseg);
            staged = pre;
            if (pre)
                fi ^= 1;
        } else {
            // ================= long track: super-blocks of caps segments chained by a carried flux =================
            const int nsb = (nseg + caps - 1) / caps;
            for (int sb = nsb - 1; sb >= 1; --sb) { // pass A: backward flux entering each super-block
                const int n = min(caps, nseg - sb * caps);
                issue_fsr(fi, cur, sb * caps, n);
                issue_ex(cur, sb * caps, n);
                gather_q(fi, cur, n);
                wait_staged();
                const int L  = ((n + 31) >> 5) | 1;
                const int lo = min(lane * L, n), hi = min(lo + L, n);
                if (lane == 0) {
#pragma unroll
                    for (int p = 0; p < P; p++)
                        sc[sb * P + p] = cb[p];
                }
                double A[P], B[P];
                chunk_compose_bwd<P>(exb, qb, lo, hi, A, B);
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
#pragma unroll
                for (int p = 0; p < P; p++)
                    cb[p] = fma(A[p], cb[p], B[p]);
                __syncwarp();
            }
            for (int sb = 0; sb < nsb; ++sb) { // pass B: forward chain, both walks, tally
                const int n = min(caps, nseg - sb * caps);
                issue_fsr(fi, cur, sb * caps, n);
                issue_ex(cur, sb * caps, n);
                gather_q(fi, cur, n);
                wait_staged();
                const int L  = ((n + 31) >> 5) | 1;
                const int lo = min(lane * L, n), hi = min(lo + L, n);
                double eb[P], psi_b[P];
#pragma unroll
                for (int p = 0; p < P; p++)
                    eb[p] = sb > 0 ? sc[sb * P + p] : cb[p];
                chunk_block<P>(exb, qb, ab, lane, lo, hi, wt, cf, eb, psi_b);
                if (sb == 0) {
#pragma unroll
                    for (int p = 0; p < P; p++)
                        cb[p] = psi_b[p];
                }
                __syncwarp();
                reduce_tally(fi, cur, n);
                __syncwarp();
            }
            staged = false;
        }

        // ---- outgoing boundary flux, written where BoundaryCondition::update would copy it ----
        if (lane == 0) {
            double *bc_out_pl = a.bc_out + (size_t)cur.plane * a.bc_per_group * GP;
#pragma unroll
            for (int p = 0; p < P; p++) {
                if (out_f[p] != INT32_MIN)
                    bc_out_pl[(size_t)(out_f[p] >= 0 ? out_f[p] : -(out_f[p] + 1)) * GP + g] = out_f[p] >= 0 ? cf[p] : 0.0;
                if (out_b[p] != INT32_MIN)
                    bc_out_pl[(size_t)(out_b[p] >= 0 ? out_b[p] : -(out_b[p] + 1)) * GP + g] = out_b[p] >= 0 ? cb[p] : 0.0;
            }
        }
        __syncwarp();
        w_cur = w_nxt, w_nxt = w_nn;
        cur = nxt;
    }
}

} // namespace mocb200
