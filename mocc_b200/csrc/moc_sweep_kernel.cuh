// moc_sweep_kernel.cuh -- the production transport-sweep kernel (sm_100a).
//
// Restates, B200-first, the inner loops of
//   sweep1g<CurrentWorker>            src/sweepers/moc/moc_sweeper_kernel.inc.hpp:84-133
//   Exponential_Linear<N>::exp        src/core/exponential.hpp:69-79
//   moc::Current::post_ray            src/sweepers/moc/moc_current_worker.hpp:202-264
//   cmdo::CurrentCorrections::post_ray  src/sweepers/cmdo/correction_worker.hpp:109-205
//   BoundaryCondition::update         src/core/boundary_condition.cpp:155-191
//
// Execution model: ONE WARP PER TRACK. A track is one ray geometry shared by the polar
// angles of a bundle; the warp sweeps it in BOTH directions for all P polar angles of the
// bundle, so every segment costs ONE red.global.add.f64 per group (forward + backward + all
// polar contributions are summed in registers first).
//
// The attenuation along a ray is an affine map per segment,
//     psi_out = a psi_in + b,   a = exp_table(-tau),  b = qbar (1 - a),
// and affine maps compose associatively. The warp therefore does not walk the ray serially:
// a lane owns 4 consecutive segments of a block of (32/GL)*4 segments, composes its 4 maps,
// warp-shuffle scans over the lanes (prefix for the forward, suffix for the backward
// direction) yield the angular flux entering every lane's chunk, and each lane then walks
// only its own 4 segments exactly like the reference loop (psi_diff = (psi - qbar) e;
// psi -= psi_diff; tally += psi_diff w). Blocks of one track are chained through a carried
// flux. Because the backward direction enters a block from the far side the kernel makes two
// passes over the track's blocks: pass 1 (last block to first) chains the backward flux and
// stores the flux entering each block; pass 2 (first to last) chains the forward flux and
// produces all tallies.
//
// GL lanes of a warp hold GL energy groups of the same segments: GL = 1 for the reference's
// per-group sweep(group) calls, 8 for group-batched sweeps (per-FSR data [n_reg][GP]
// group-fastest: one coalesced 64-byte access per segment).
//
// CACHED = true (production when it fits): a = exp_table(-xstr*len/sin(theta)) only changes
// when the host uploads new cross sections, while a sweep(group) call runs n_inner inner
// iterations of two passes each. exp_cache_kernel evaluates the table ONCE per upload (same
// shared-memory table, same interpolation: bit-identical values) into an HBM-resident array
// and the sweep streams 8 bytes per (segment, polar angle, group) -- a coalesced stream
// B200's HBM3e delivers faster than the SMs can redo the lookups (2 bank-conflicted LDS +
// ~8 FP64 ops each). The sweep then reads neither lengths nor cross sections nor the table.
// For GL = 1 the scattered q-bar gathers and tally reductions are issued with the lanes on 32
// CONSECUTIVE segments ("striped": neighbouring segments lie in the same pin, hence in the
// same few 128-byte lines of the group-major q/tally arrays) and moved to/from the
// lane-owns-4-consecutive-segments ("blocked") arrangement by a padded, conflict-free
// shared-memory transpose.
// CACHED = false (fallback): the table is staged into shared memory with a TMA bulk copy
// (cp.async.bulk + mbarrier) and evaluated in both passes.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "moc_kernels.cuh"

namespace mocb200 {

// One (track, polar bundle). 32 bytes.
struct __align__(16) TrackUnit {
    int32_t seg_begin; // first segment in the PADDED segment arrays (multiple of 4)
    int32_t nseg;
    int32_t bc0; // Ray::bc(0): forward entry slot / backward exit slot
    int32_t bc1; // Ray::bc(1): forward exit slot / backward entry slot
    int32_t bundle;
    int32_t cpos; // position of the unit's first (padded) segment inside its list's attenuation cache
    int32_t pad1, pad2;
};

struct WarpArgs {
    const TrackUnit *units;
    int32_t n_units;
    uint32_t *counter;
    const Bundle *bundles;
    const int32_t *planes;
    int32_t n_planes;
    // padded geometry
    const double *seg_len;
    const int32_t *seg_fsr;
    const int2 *xptr;   // per 4 segments: first fwd / bwd crossing index (TALLY)
    const Cross *cross; // crossing lists with sentinels
    // angle tables
    const double *ang_rsintheta;
    const double *wt_v_st; // [n_plane][n_ang]
    const double *cur_w;   // [n_plane][n_ang][2]
    const double *flx_w;
    const int32_t *bc_offset;
    const int32_t *bc_size_x;
    const int32_t *bc_dst_off;
    const int32_t *bc_dst_kind;
    const int32_t *plane_first_reg;
    const int32_t *plane_surf_offset;
    int32_t n_ang, bc_per_group, n_plane_total, n_surf_plane;
    // group data
    int32_t g_begin, g_count, GP, n_gsets, n_reg;
    const double2 *xq; // !CACHED: [n_reg][GP] {xstr, qbar}
    const double *q;   // CACHED: GL = 1 group-major [g][n_reg]; GL = 8 [n_reg][GP]
    double *tally;     // same layout as q (CACHED) / [n_reg][GP] (!CACHED)
    const double *bc_in;
    double *bc_out;
    double *current;      // [n_surf][GP]
    double *surface_flux; // [n_surf][GP]
    // TALLY == 2 (cmdo::CurrentCorrections): per-angle, per-direction sums
    double *dsum; // psi_diff per FSR:  GL=1 [g][n_reg][2 n_ang]; GL=8 [n_reg][2 n_ang][GP]
    double *ssum; // psi per crossing:  GL=1 [g][plane][n_ang][n_surf_plane][2]; GL=8 [plane][n_ang][n_surf_plane][2][GP]
    // per-warp scratch: flux entering each block in the backward direction
    double *scratch;
    int32_t scratch_per_warp;
    // attenuation cache of this list: GL = 1 [plane][g][pos][P]; GL = 8 [plane][pos][P][GP]
    const double *cache;
    int64_t list_pseg;
    int32_t cache_groups, cache_g0; // groups per plane in the cache, first group it holds
    // exponential table (!CACHED)
    const double *exp_table;
    int32_t exp_n;
    double exp_min, exp_max;
};

// ---- TMA (bulk async copy) staging of the exponential table ----
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(phase)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Linear interpolation in the reference's table (exponential.hpp:69-79): same grid, same
// table entries, same interpolant. The interval index is obtained with a round-down add
// instead of a double->int conversion and the interpolation weight as x - floor(x); both
// differ from the reference's operation order by O(1e-15) relative (continuity at the
// knots makes an index flip at an interval boundary harmless). Arguments outside the table
// [vmin, vmax] (x < 0 or x > xmax = N; a non-positive cross section under transverse-leakage
// splitting gives v > 0) fall back to exp() as the reference does (exponential.hpp:71-75, minus its print).
__device__ __forceinline__ double exp_interp(const double *__restrict__ tab, double v, double c0, double rspace,
                                             double xmax)
{
    const double x = fma(v, rspace, c0); // (v - vmin) * rspace
    if (x < 0.0 || x > xmax)
        return exp(v);
    const double magic = 4503599627370496.0; // 2^52
    const double xi    = __dadd_rd(x, magic);
    const int i        = __double2loint(xi);
    const double frac  = x - (xi - magic);
    const double d0    = tab[i];
    const double d1    = tab[i + 1];
    return fma(d1 - d0, frac, d0);
}

constexpr int kWarpBlock        = 512; // threads per CTA (16 warps), one persistent CTA per SM
constexpr int kTransposeDoubles = 144; // 128 + 2 per 16: conflict-free 16-byte blocked accesses

__device__ __forceinline__ int tpos(int s)
{
    return s + 2 * (s >> 4);
}

template <int GL, int P, int TALLY, bool CACHED>
static __global__ void __launch_bounds__(kWarpBlock, 1) sweep_warp_kernel(const WarpArgs a)
{
    constexpr int C       = 4;
    constexpr int NCH     = 32 / GL; // chunk lanes per warp
    constexpr int SEGB    = NCH * C; // segments per block
    constexpr bool STRIPE = CACHED && GL == 1;

    extern __shared__ __align__(16) double s_dyn[]; // !CACHED: exponential table
    __shared__ __align__(16) double s_tr[STRIPE ? (kWarpBlock / 32) * kTransposeDoubles : 2];
    __shared__ uint64_t s_bar;

    double c0 = 0.0, rspace = 0.0;
    const double xmax = (double)a.exp_n;
    if (!CACHED) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = ((uint32_t)(a.exp_n + 2) * 8u + 15u) & ~15u;
            mbar_expect_tx(&s_bar, bytes);
            for (uint32_t off = 0; off < bytes; off += 32768u) { // <= 32 KB per bulk copy
                const uint32_t n = min(32768u, bytes - off);
                bulk_g2s(reinterpret_cast<char *>(s_dyn) + off, reinterpret_cast<const char *>(a.exp_table) + off, n,
                         &s_bar);
            }
        }
        mbar_wait(&s_bar, 0);
        const double space = (a.exp_max - a.exp_min) / (double)a.exp_n;
        rspace             = 1.0 / space;
        c0                 = -a.exp_min * rspace;
    }

    const int lane = threadIdx.x & 31;
    const int ch   = lane / GL;
    const int gl   = lane - ch * GL;
    const int GP   = a.GP;
    double *tr     = s_tr + (STRIPE ? (threadIdx.x >> 5) * kTransposeDoubles : 0);

    const uint32_t per_unit = (uint32_t)a.n_planes * (uint32_t)a.n_gsets;
    const uint32_t total    = (uint32_t)a.n_units * per_unit;
    const int warp_global   = (blockIdx.x * (kWarpBlock / 32)) + (threadIdx.x >> 5);
    double *sc              = a.scratch + (size_t)warp_global * a.scratch_per_warp;
    const int32_t *__restrict__ seg_fsr = a.seg_fsr;
    const double *__restrict__ seg_len  = a.seg_len;
    const int nslot                     = 2 * a.n_ang;

    for (;;) {
        uint32_t w = 0;
        if (lane == 0)
            w = atomicAdd(a.counter, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= total)
            break;
        const int unit_id = (int)(w / per_unit);
        const uint32_t r  = w - (uint32_t)unit_id * per_unit;
        const int ipl     = (int)(r / (uint32_t)a.n_gsets);
        const int gset    = (int)(r - (uint32_t)ipl * (uint32_t)a.n_gsets);
        const int plane   = a.planes[ipl];
        const int first_reg = a.plane_first_reg[plane];

        const int4 u0 = reinterpret_cast<const int4 *>(a.units)[2 * unit_id];
        const int4 u1 = reinterpret_cast<const int4 *>(a.units)[2 * unit_id + 1];
        const int seg_begin = u0.x, nseg = u0.y, bc0 = u0.z, bc1 = u0.w;
        const int bundle = u1.x, cpos = u1.y;
        const int npad = (nseg + 3) & ~3;

        int g          = a.g_begin + gset * GL + gl;
        const bool gok = g < a.g_begin + a.g_count;
        if (!gok)
            g = a.g_begin; // idle group lane: computes on valid addresses, never writes
        const int grel = g - a.g_begin;

        // per-FSR arrays and attenuation stream of this (plane, group)
        const double *__restrict__ qv   = nullptr;
        double *__restrict__ tv         = nullptr;
        const double *__restrict__ ex_b = nullptr;
        size_t fstride                  = 1; // stride between FSRs in qv / tv
        if (CACHED) {
            if (GL == 1) {
                qv   = a.q + (size_t)grel * a.n_reg;
                tv   = a.tally + (size_t)grel * a.n_reg;
                ex_b = a.cache + (((size_t)ipl * a.cache_groups + (g - a.cache_g0)) * a.list_pseg + cpos) * P;
            } else {
                qv      = a.q + g;
                tv      = a.tally + g;
                fstride = GP;
                ex_b    = a.cache + (((size_t)ipl * a.list_pseg + cpos) * P) * GP + g;
            }
        } else {
            tv      = a.tally + g;
            fstride = GP;
        }

        double wt[P], nrs[P], cf[P], cb[P];
        int ang[P];
        const double *bc_in_pl = a.bc_in + (size_t)plane * a.bc_per_group * GP;
#pragma unroll
        for (int p = 0; p < P; p++) {
            ang[p] = a.bundles[bundle].ang[p];
            wt[p]  = a.wt_v_st[plane * a.n_ang + ang[p]];
            nrs[p] = CACHED ? 0.0 : -a.ang_rsintheta[ang[p]];
            cf[p]  = bc_in_pl[(size_t)(a.bc_offset[ang[p]] + bc0) * GP + g];
            cb[p]  = bc_in_pl[(size_t)(a.bc_offset[ang[p] + a.n_ang] + bc1) * GP + g];
        }
        double cw[P][2], fw[P][2];
        int surf_off = 0;
        if (TALLY != 0) {
            surf_off = a.plane_surf_offset[plane];
#pragma unroll
            for (int p = 0; p < P; p++) {
                const size_t o = ((size_t)plane * a.n_ang + ang[p]) * 2;
                cw[p][0] = a.cur_w[o], cw[p][1] = a.cur_w[o + 1];
                fw[p][0] = a.flx_w[o], fw[p][1] = a.flx_w[o + 1];
            }
        }
        const int nblk = (nseg + SEGB - 1) / SEGB;

        // Loads one block: q-bar and attenuations in the BLOCKED arrangement (index c = the lane's own
        // c-th segment); FSR ids in fs[] are STRIPED when STRIPE (used for the tally reduction) else blocked.
        auto load_block = [&](int b, double (&q)[C], double (&ex)[P][C], int (&fs)[C]) {
            const int k0 = b * SEGB + ch * C;
            if (STRIPE) {
                double qs[C];
#pragma unroll
                for (int i = 0; i < C; i++) {
                    const int k = b * SEGB + 32 * i + lane;
                    fs[i]       = k < npad ? seg_fsr[seg_begin + k] + first_reg : first_reg;
                    qs[i]       = qv[fs[i]];
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < C; i++)
                    tr[tpos(32 * i + lane)] = qs[i];
                __syncwarp();
                const double2 q01 = *reinterpret_cast<const double2 *>(tr + tpos(4 * lane));
                const double2 q23 = *reinterpret_cast<const double2 *>(tr + tpos(4 * lane) + 2);
                q[0] = q01.x, q[1] = q01.y, q[2] = q23.x, q[3] = q23.y;
            } else {
                if (k0 < npad) {
                    const int4 f = *reinterpret_cast<const int4 *>(seg_fsr + seg_begin + k0);
                    fs[0] = f.x + first_reg, fs[1] = f.y + first_reg, fs[2] = f.z + first_reg, fs[3] = f.w + first_reg;
                } else {
#pragma unroll
                    for (int c = 0; c < C; c++)
                        fs[c] = first_reg;
                }
            }
            if (CACHED) {
                if (!STRIPE) {
#pragma unroll
                    for (int c = 0; c < C; c++)
                        q[c] = qv[(size_t)fs[c] * fstride];
                }
                if (k0 < npad) {
                    if (GL == 1) {
                        const double2 *src = reinterpret_cast<const double2 *>(ex_b + (size_t)k0 * P);
                        double buf[C * P];
#pragma unroll
                        for (int j = 0; j < C * P / 2; j++) {
                            const double2 t = src[j];
                            buf[2 * j] = t.x, buf[2 * j + 1] = t.y;
                        }
#pragma unroll
                        for (int c = 0; c < C; c++)
#pragma unroll
                            for (int p = 0; p < P; p++)
                                ex[p][c] = buf[c * P + p];
                    } else {
#pragma unroll
                        for (int c = 0; c < C; c++)
#pragma unroll
                            for (int p = 0; p < P; p++)
                                ex[p][c] = ex_b[((size_t)(k0 + c) * P + p) * GP];
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < C; c++)
#pragma unroll
                        for (int p = 0; p < P; p++)
                            ex[p][c] = 1.0;
                }
            } else {
                double len[C];
                if (k0 < npad) {
                    const double2 l01 = *reinterpret_cast<const double2 *>(seg_len + seg_begin + k0);
                    const double2 l23 = *reinterpret_cast<const double2 *>(seg_len + seg_begin + k0 + 2);
                    len[0] = l01.x, len[1] = l01.y, len[2] = l23.x, len[3] = l23.y;
                } else {
#pragma unroll
                    for (int c = 0; c < C; c++)
                        len[c] = 0.0;
                }
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const double2 v  = a.xq[(size_t)fs[c] * GP + g];
                    const bool valid = k0 + c < nseg;
                    const double t   = v.x * len[c];
                    q[c]             = v.y;
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        const double x = exp_interp(s_dyn, t * nrs[p], c0, rspace, xmax);
                        ex[p][c]       = valid ? x : 1.0;
                    }
                }
            }
        };

        // ================= pass 1: backward flux entering each block =================
        for (int b = nblk - 1; b >= 0; --b) {
            double q[C], ex[P][C];
            int fs[C];
            load_block(b, q, ex, fs);
            if (ch == 0 && gok) {
#pragma unroll
                for (int p = 0; p < P; p++)
                    sc[(b * P + p) * GL + gl] = cb[p];
            }
            double A[P], B[P];
#pragma unroll
            for (int p = 0; p < P; p++)
                A[p] = 1.0, B[p] = 0.0;
#pragma unroll
            for (int c = 0; c < C; c++) {
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const double bq = q[c] * (1.0 - ex[p][c]);
                    B[p] = fma(A[p], bq, B[p]); // M o m_c: the backward sweep applies the higher segment first
                    A[p] *= ex[p][c];
                }
            }
            // ordered butterfly reduction over the chunk lanes: total = L_0 o L_1 o ... o L_{NCH-1}
#pragma unroll
            for (int s = GL; s < 32; s <<= 1) {
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
        }

        // ================= pass 2: forward chain, all tallies =================
        for (int b = 0; b < nblk; ++b) {
            const int k0 = b * SEGB + ch * C;
            double q[C], ex[P][C];
            int fs[C];
            load_block(b, q, ex, fs);
            int2 xp = make_int2(0, 0);
            if (TALLY != 0 && k0 < nseg)
                xp = a.xptr[(seg_begin + k0) >> 2];
            int fb[C]; // blocked FSR ids (per-FSR correction sums)
            if (TALLY == 2) {
                if (STRIPE) {
                    int4 f = make_int4(0, 0, 0, 0);
                    if (k0 < npad)
                        f = *reinterpret_cast<const int4 *>(seg_fsr + seg_begin + k0);
                    fb[0] = f.x + first_reg, fb[1] = f.y + first_reg, fb[2] = f.z + first_reg, fb[3] = f.w + first_reg;
                } else {
#pragma unroll
                    for (int c = 0; c < C; c++)
                        fb[c] = fs[c];
                }
            }
            // flux entering this block in the backward direction (written by this very lane in pass 1)
            double eb[P];
#pragma unroll
            for (int p = 0; p < P; p++) {
                double x = 0.0;
                if (ch == 0 && gok)
                    x = sc[(b * P + p) * GL + gl];
                eb[p] = __shfl_sync(0xffffffffu, x, gl);
            }
            double A[P], Bf[P], Bb[P];
#pragma unroll
            for (int p = 0; p < P; p++)
                A[p] = 1.0, Bf[p] = 0.0, Bb[p] = 0.0;
#pragma unroll
            for (int c = 0; c < C; c++) {
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const double bq = q[c] * (1.0 - ex[p][c]);
                    Bb[p] = fma(A[p], bq, Bb[p]);     // M o m_c
                    Bf[p] = fma(ex[p][c], Bf[p], bq); // m_c o M
                    A[p] *= ex[p][c];
                }
            }
            // inclusive scans over the chunk lanes: prefix for the forward, suffix for the backward direction
            double psi_f[P], psi_b[P];
            {
                double Af[P], Ab[P];
#pragma unroll
                for (int p = 0; p < P; p++)
                    Af[p] = A[p], Ab[p] = A[p];
#pragma unroll
                for (int s = 1; s < NCH; s <<= 1) {
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        const double Ae = __shfl_up_sync(0xffffffffu, Af[p], s * GL);
                        const double Be = __shfl_up_sync(0xffffffffu, Bf[p], s * GL);
                        const double Ah = __shfl_down_sync(0xffffffffu, Ab[p], s * GL);
                        const double Bh = __shfl_down_sync(0xffffffffu, Bb[p], s * GL);
                        if (ch >= s) { // mine o earlier
                            Bf[p] = fma(Af[p], Be, Bf[p]);
                            Af[p] *= Ae;
                        }
                        if (ch + s < NCH) { // mine o higher
                            Bb[p] = fma(Ab[p], Bh, Bb[p]);
                            Ab[p] *= Ah;
                        }
                    }
                }
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const double out_f = fma(Af[p], cf[p], Bf[p]); // flux leaving this lane's chunk, forward
                    const double out_b = fma(Ab[p], eb[p], Bb[p]); // ... backward
                    const double in_f  = __shfl_up_sync(0xffffffffu, out_f, GL);
                    const double in_b  = __shfl_down_sync(0xffffffffu, out_b, GL);
                    psi_f[p] = ch == 0 ? cf[p] : in_f;
                    psi_b[p] = ch == NCH - 1 ? eb[p] : in_b;
                    cf[p]    = __shfl_sync(0xffffffffu, out_f, (NCH - 1) * GL + gl); // carried to the next block
                }
            }

            // ---- walk the lane's own 4 segments like the reference loop ----
            double acc[C];
            Cross xf, xb;
            int ci_f = xp.x, ci_b = xp.y;
            if (TALLY != 0) {
                xf = a.cross[ci_f];
                xb = a.cross[ci_b];
            }
            // coarse-surface crossing: moc::Current (TALLY 1) or cmdo::CurrentCorrections (TALLY 2)
            auto tally_cross = [&](const Cross &x, const double (&psi)[P], int dir) {
                const int norm  = x.surf & 1;
                const int surf  = x.surf >> 1;
                const size_t o  = (size_t)(surf + surf_off) * GP + g;
                double cs = 0.0, fsum = 0.0;
#pragma unroll
                for (int p = 0; p < P; p++) {
                    cs   = fma(psi[p], cw[p][norm], cs);
                    fsum = fma(psi[p], fw[p][norm], fsum);
                }
                // forward adds, backward subtracts (moc_current_worker.hpp:230-231); the corrections
                // worker also subtracts the backward SURFACE FLUX (correction_worker.hpp:136-137, 194-195)
                atomicAdd(&a.current[o], dir ? -cs : cs);
                atomicAdd(&a.surface_flux[o], (dir && TALLY == 2) ? -fsum : fsum);
                if (TALLY == 2) {
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        size_t so = (((size_t)plane * a.n_ang + ang[p]) * a.n_surf_plane + surf) * 2 + dir;
                        if (GL == 1)
                            so += (size_t)grel * a.n_plane_total * a.n_ang * a.n_surf_plane * 2;
                        else
                            so = so * GP + g;
                        atomicAdd(&a.ssum[so], psi[p]);
                    }
                }
            };
            auto dsum_add = [&](int reg, int p, int dir, double d) {
                size_t o = (size_t)reg * nslot + ang[p] * 2 + dir;
                if (GL == 1)
                    o += (size_t)grel * a.n_reg * nslot;
                else
                    o = o * GP + g;
                atomicAdd(&a.dsum[o], d);
            };
#pragma unroll
            for (int c = 0; c < C; c++) {
                if (TALLY != 0 && gok) {
                    const int node = k0 + c; // forward flux at the node in front of segment k0+c
                    while (xf.node == node && node < nseg) {
                        tally_cross(xf, psi_f, 0);
                        xf = a.cross[++ci_f];
                    }
                }
                double s = 0.0;
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const double d = (psi_f[p] - q[c]) * (1.0 - ex[p][c]);
                    psi_f[p] -= d;
                    s = fma(d, wt[p], s);
                    if (TALLY == 2 && gok && k0 + c < nseg)
                        dsum_add(fb[c], p, 0, d);
                }
                acc[c] = s;
                if (TALLY != 0 && gok && k0 + c == nseg - 1) { // far end of the ray
                    while (xf.node == nseg) {
                        tally_cross(xf, psi_f, 0);
                        xf = a.cross[++ci_f];
                    }
                }
            }
#pragma unroll
            for (int c = C - 1; c >= 0; c--) {
                const int k = k0 + c;
                if (TALLY != 0 && gok && k < nseg) {
                    const int nb = nseg - 1 - k; // segments walked by the backward sweep so far
                    while (xb.node == nb) {
                        tally_cross(xb, psi_b, 1);
                        xb = a.cross[++ci_b];
                    }
                }
                double s = acc[c];
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const double d = (psi_b[p] - q[c]) * (1.0 - ex[p][c]);
                    psi_b[p] -= d;
                    s = fma(d, wt[p], s);
                    if (TALLY == 2 && gok && k < nseg)
                        dsum_add(fb[c], p, 1, d);
                }
                acc[c] = s;
                if (TALLY != 0 && gok && k == 0) { // near end of the ray
                    while (xb.node == nseg) {
                        tally_cross(xb, psi_b, 1);
                        xb = a.cross[++ci_b];
                    }
                }
            }
            // ---- scalar-flux tally: one reduction per segment and group ----
            if (STRIPE) {
                __syncwarp();
                *reinterpret_cast<double2 *>(tr + tpos(4 * lane))     = make_double2(acc[0], acc[1]);
                *reinterpret_cast<double2 *>(tr + tpos(4 * lane) + 2) = make_double2(acc[2], acc[3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < C; i++) {
                    const int k = b * SEGB + 32 * i + lane;
                    if (k < nseg)
                        atomicAdd(&tv[fs[i]], tr[tpos(32 * i + lane)]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < C; c++)
                    if (gok && k0 + c < nseg)
                        atomicAdd(&tv[(size_t)fs[c] * fstride], acc[c]);
            }
            if (b == 0) { // backward flux leaving the ray: lane of the first chunk of the first block
#pragma unroll
                for (int p = 0; p < P; p++)
                    cb[p] = psi_b[p];
            }
        }

        // ---- outgoing boundary flux, written where BoundaryCondition::update would copy it ----
        if (ch == 0 && gok) {
            double *bc_out_pl = a.bc_out + (size_t)plane * a.bc_per_group * GP;
#pragma unroll
            for (int p = 0; p < P; p++) {
#pragma unroll
                for (int dir = 0; dir < 2; dir++) {
                    const int ao       = ang[p] + dir * a.n_ang;
                    const int out_slot = dir ? bc0 : bc1;
                    const double psi   = dir ? cb[p] : cf[p];
                    const int sx       = a.bc_size_x[ao];
                    const int face     = out_slot >= sx ? 1 : 0;
                    const int idx      = out_slot - (face ? sx : 0);
                    const int kind     = a.bc_dst_kind[2 * ao + face];
                    if (kind != 2) {
                        const size_t o = (size_t)(a.bc_dst_off[2 * ao + face] + idx) * GP + g;
                        bc_out_pl[o]   = (kind == 1) ? psi : 0.0;
                    }
                }
            }
        }
    }
}

// Fills the attenuation cache of one list for groups [g_begin, g_begin + g_count): the
// reference's table lookup (exponential.hpp:69-79) of -xstr*len/sin(theta), evaluated once per
// cross-section upload. One warp per (unit, plane); padded segments get the identity (1.0).
struct CacheArgs {
    const TrackUnit *units;
    int32_t n_units;
    const Bundle *bundles;
    const int32_t *planes;
    int32_t n_planes;
    const double *seg_len;
    const int32_t *seg_fsr;
    const int4 *len_begin; // per unit and polar angle: start of that angle's own (padded) segment lengths
    const double *ang_rsintheta;
    const int32_t *plane_first_reg;
    const double *xstr; // [n_reg][GP]
    int32_t g_begin, g_count, cache_groups, cache_g0, GP, np, group_major;
    double *cache;
    int64_t list_pseg;
    const double *exp_table;
    int32_t exp_n;
    double exp_min, exp_max;
};

template <int P> __global__ void __launch_bounds__(512, 1) exp_cache_kernel(const CacheArgs a)
{
    extern __shared__ __align__(16) double s_tab[];
    for (int i = threadIdx.x; i < a.exp_n + 2; i += blockDim.x)
        s_tab[i] = a.exp_table[i];
    __syncthreads();
    const double space  = (a.exp_max - a.exp_min) / (double)a.exp_n;
    const double rspace = 1.0 / space;
    const double c0     = -a.exp_min * rspace;
    const double xmax   = (double)a.exp_n;
    const int lane      = threadIdx.x & 31;
    const int warps     = gridDim.x * (blockDim.x >> 5);
    const int64_t total = (int64_t)a.n_units * a.n_planes;
    for (int64_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += warps) {
        const int unit_id   = (int)(w / a.n_planes);
        const int ipl       = (int)(w - (int64_t)unit_id * a.n_planes);
        const int plane     = a.planes[ipl];
        const int first_reg = a.plane_first_reg[plane];
        const TrackUnit u   = a.units[unit_id];
        const int4 lb4      = a.len_begin[unit_id];
        const int lb[4]     = {lb4.x, lb4.y, lb4.z, lb4.w};
        double nrs[P];
#pragma unroll
        for (int p = 0; p < P; p++)
            nrs[p] = -a.ang_rsintheta[a.bundles[u.bundle].ang[p]];
        const int npad = (u.nseg + 3) & ~3;
        for (int k = lane; k < npad; k += 32) {
            const bool valid = k < u.nseg;
            const int reg    = a.seg_fsr[u.seg_begin + k] + first_reg;
            double len[P]; // every polar angle has its own lengths (polar copies may differ in the last bits)
#pragma unroll
            for (int p = 0; p < P; p++)
                len[p] = a.seg_len[lb[p] + k];
            for (int gi = 0; gi < a.g_count; gi++) {
                const int g     = a.g_begin + gi;
                const double xs = a.xstr[(size_t)reg * a.GP + g];
                double ex[P];
#pragma unroll
                for (int p = 0; p < P; p++)
                    ex[p] = valid ? exp_interp(s_tab, xs * len[p] * nrs[p], c0, rspace, xmax) : 1.0;
                if (a.group_major) { // [plane][g][pos][P]: the P values of a position are contiguous
                    double *dst = a.cache + (((size_t)ipl * a.cache_groups + (g - a.cache_g0)) * a.list_pseg + u.cpos + k) * P;
                    if constexpr (P == 2) {
                        *reinterpret_cast<double2 *>(dst) = make_double2(ex[0], ex[1]);
                    } else if constexpr (P == 4) {
                        *reinterpret_cast<double2 *>(dst)     = make_double2(ex[0], ex[1]);
                        *reinterpret_cast<double2 *>(dst + 2) = make_double2(ex[2], ex[3]);
                    } else {
#pragma unroll
                        for (int p = 0; p < P; p++)
                            dst[p] = ex[p];
                    }
                } else { // [plane][pos][P][GP]
#pragma unroll
                    for (int p = 0; p < P; p++)
                        a.cache[(((size_t)ipl * a.list_pseg + u.cpos + k) * P + p) * a.GP + g] = ex[p];
                }
            }
        }
    }
}

// q-bar = (src + flux*xs_self) * (1/(xstr_src*4pi)) (source_isotropic.cpp:29-31, non-contracted
// arithmetic) into every layout the sweep kernels read, plus the tally reset.
//   group_major: q_out/tally_out are [g - g_begin][n_reg], else [n_reg][GP]; xq gets {xstr, q}.
static __global__ void self_scatter_q_kernel(int n_reg, int GP, int g_begin, int g_count, const double *__restrict__ src,
                                      const double *__restrict__ flux, const double *__restrict__ xs_self,
                                      const double *__restrict__ xstr_src, const double *__restrict__ xstr,
                                      double *qbar, double *q_out, double2 *__restrict__ xq,
                                      double *__restrict__ tally_out, int group_major, int compute_q,
                                      uint32_t *__restrict__ counters, int n_counters,
                                      const int32_t *__restrict__ perm = nullptr, int n_regp = 0)
{
    if (blockIdx.x == 0 && (int)threadIdx.x < n_counters) // work counters of the sweep kernels that follow
        counters[threadIdx.x] = 0u;
    const int64_t n = (int64_t)n_reg * g_count;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int r, gi;
        if (group_major) {
            gi = (int)(i / n_reg);
            r  = (int)(i - (int64_t)gi * n_reg);
        } else {
            r  = (int)(i / g_count);
            gi = (int)(i - (int64_t)r * g_count);
        }
        const int g    = g_begin + gi;
        const size_t o = (size_t)r * GP + g;
        double q;
        if (compute_q) {
            const double r_fpi_tr = __ddiv_rn(1.0, __dmul_rn(xstr_src[o], kFPi));
            q = __dmul_rn(__dadd_rn(src[o], __dmul_rn(flux[o], xs_self[o])), r_fpi_tr);
            qbar[o] = q;
        } else {
            q = qbar[o];
        }
        // group-major layout of the register-chunk kernel: FSRs regrouped so that FSRs visited together share 32-byte sectors
        const size_t oo = group_major ? (perm ? (size_t)gi * n_regp + perm[r] : (size_t)gi * n_reg + r) : o;
        if (xq)
            xq[o] = make_double2(xstr[o], q);
        else
            q_out[oo] = q;
        tally_out[oo] = 0.0;
    }
}

// flux = tally/(xstr*vol) + qbar*4pi   (kernel:165-173), tally in either layout
static __global__ void finalize_flux_q_kernel(int n_reg, int GP, int g_begin, int g_count, const double *__restrict__ tally,
                                       const double *__restrict__ xstr, const double *__restrict__ vol,
                                       const double *__restrict__ qbar, double *__restrict__ flux, int reg_lo,
                                       int reg_hi, int group_major, const int32_t *__restrict__ perm = nullptr,
                                       int n_regp = 0)
{
    const int nr    = reg_hi - reg_lo;
    const int64_t n = (int64_t)nr * g_count;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int r, gi;
        if (group_major) {
            gi = (int)(i / nr);
            r  = reg_lo + (int)(i - (int64_t)gi * nr);
        } else {
            r  = reg_lo + (int)(i / g_count);
            gi = (int)(i % g_count);
        }
        const int g     = g_begin + gi;
        const size_t o  = (size_t)r * GP + g;
        const size_t oo = group_major ? (perm ? (size_t)gi * n_regp + perm[r] : (size_t)gi * n_reg + r) : o;
        flux[o] = __dadd_rn(__ddiv_rn(tally[oo], __dmul_rn(xstr[o], vol[r])), __dmul_rn(qbar[o], kFPi));
    }
}

// finalize_flux_q_kernel of inner iteration i fused with self_scatter_q_kernel of iteration i + 1 (group-major
// q-bar / tally of the per-group sweeps): flux = tally/(xstr*vol) + qbar*4pi; qbar' = (src + flux*xs_self) /
// (xstr_src*4pi); tally = 0; work counters = 0. Same non-contracted arithmetic as the two kernels.
static __global__ void finalize_next_q_kernel(int n_reg, int GP, int g_begin, int g_count, double *__restrict__ tally,
                                       const double *__restrict__ xstr, const double *__restrict__ vol,
                                       double *__restrict__ qbar, double *__restrict__ flux, int reg_lo, int reg_hi,
                                       const double *__restrict__ src, const double *__restrict__ xs_self,
                                       const double *__restrict__ xstr_src, double *__restrict__ q_out,
                                       uint32_t *__restrict__ counters, int n_counters,
                                       const int32_t *__restrict__ perm = nullptr, int n_regp = 0)
{
    if (blockIdx.x == 0 && (int)threadIdx.x < n_counters)
        counters[threadIdx.x] = 0u;
    const int nr    = reg_hi - reg_lo;
    const int64_t n = (int64_t)nr * g_count;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int gi    = (int)(i / nr);
        const int r     = reg_lo + (int)(i - (int64_t)gi * nr);
        const int g     = g_begin + gi;
        const size_t o  = (size_t)r * GP + g;
        const size_t oo = perm ? (size_t)gi * n_regp + perm[r] : (size_t)gi * n_reg + r;
        const double f  = __dadd_rn(__ddiv_rn(tally[oo], __dmul_rn(xstr[o], vol[r])), __dmul_rn(qbar[o], kFPi));
        flux[o]         = f;
        const double r_fpi_tr = __ddiv_rn(1.0, __dmul_rn(xstr_src[o], kFPi));
        const double q        = __dmul_rn(__dadd_rn(src[o], __dmul_rn(f, xs_self[o])), r_fpi_tr);
        qbar[o]   = q;
        q_out[oo] = q;
        tally[oo] = 0.0;
    }
}

// ---- source construction on the device ----
// Cross sections by cross-section-mesh region (material): fsr_mat [n_reg], xsnf / xsch [n_mat][G],
// scat [n_mat][G to][G from] with the band [lo, hi] of every row (ScatteringRow::min_g / max_g).
// TransportSweeper::calc_fission_source (transport_sweeper.cpp:119-134): fs = 0; += (1/k * nu-Sigma_f,g) * flux_g, g ascending
static __global__ void fission_source_kernel(int reg_lo, int reg_hi, int G, int GP, double rkeff,
                                             const int32_t *__restrict__ fsr_mat, const double *__restrict__ xsnf,
                                             const double *__restrict__ flux, double *__restrict__ fs)
{
    for (int r = reg_lo + blockIdx.x * blockDim.x + threadIdx.x; r < reg_hi; r += gridDim.x * blockDim.x) {
        const double *nf = xsnf + (size_t)fsr_mat[r] * G;
        double acc       = 0.0;
        for (int g = 0; g < G; g++)
            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(rkeff, nf[g]), flux[(size_t)r * GP + g]));
        fs[r] = acc;
    }
}

// What FixedSourceSolver::step builds before sweep(group) (fixed_source_solver.cpp:102-117): Source::initialize_group
// (source.cpp:41-54: external source or 0) + Source::fission (:64-80: += chi_g fs) + Source::in_scatter (:86-112: the
// row's band in ascending order without the self-scatter term, += Sigma_s(g' -> g) flux_g'). Reads the RESIDENT flux: the
// groups already swept in this outer carry their new flux, as in the reference's Gauss-Seidel over groups.
static __global__ void group_source_kernel(int reg_lo, int reg_hi, int G, int GP, int g_begin, int g_count,
                                           const int32_t *__restrict__ fsr_mat, const double *__restrict__ xsch,
                                           const double *__restrict__ scat, const int32_t *__restrict__ band,
                                           const double *__restrict__ ext, const double *__restrict__ fs,
                                           const double *__restrict__ flux, double *__restrict__ src)
{
    const int nr    = reg_hi - reg_lo;
    const int64_t n = (int64_t)nr * g_count;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = reg_lo + (int)(i / g_count);
        const int g = g_begin + (int)(i % g_count);
        const int m = fsr_mat[r];
        double s_   = ext ? ext[(size_t)r * GP + g] : 0.0;
        s_          = __dadd_rn(s_, __dmul_rn(xsch[(size_t)m * G + g], fs[r]));
        const double *row = scat + ((size_t)m * G + g) * G;
        const int lo = band[((size_t)m * G + g) * 2], hi = band[((size_t)m * G + g) * 2 + 1];
        for (int gg = lo; gg <= hi; gg++)
            if (gg != g)
                s_ = __dadd_rn(s_, __dmul_rn(row[gg], flux[(size_t)r * GP + gg]));
        src[(size_t)r * GP + g] = s_;
    }
}

// cmdo::CurrentCorrections::post_angle + calculate_corrections (correction_worker.hpp:223-246,
// correction_worker.cpp:32-158) from the per-angle sums the TALLY == 2 sweep left behind.
// One thread per (plane of this handle, sweep angle, coarse cell, direction).
//   vol_sum  = sum over the cell's segments of  t*qbar + psi_diff/xstr_split
//            = sum over the cell's FSRs of      T*qbar + D/xstr_split      (T = sum of t, D = sum of psi_diff)
//   sigt_sum = same with every term times xstr_true;  vol_norm = sum of T
struct CorrArgs {
    const int32_t *planes; // macroplanes of this handle
    int32_t n_planes, n_ang, n_cell_plane, n_surf_plane, n_plane_total, n_geom, GP, n_reg, gl;
    int32_t g_begin, g_count;
    const int32_t *plane_unique, *plane_first_reg, *plane_cell_offset;
    const int32_t *uniq_reg_begin; // [n_unique + 1] offsets into plane-local FSR tables
    const int32_t *cell_fsr_begin; // [n_unique][n_cell_plane + 1] CSR (offsets relative to the unique plane's list)
    const int32_t *cell_fsr;       // plane-local FSR ids, cell by cell, per unique plane (same offsets as uniq_reg_begin)
    const double *geom_len;        // [n_geom] blocks at uniq_reg_begin[u]*n_geom + geom*nreg_u : path length per FSR
    const int32_t *ang_geom;
    const double *ang_rsintheta, *ang_area_x, *ang_area_y, *ang_ox;
    const double *cell_dx, *cell_dy;
    const int32_t *coarse_surf; // [n_cell_plane][4] E,N,W,S
    const double *xstr, *xstr_true, *qbar; // [n_reg][GP]
    const double *sn_xs;                   // [n_plane_total*n_cell_plane][GP]
    const double *dsum, *ssum;
    double *alpha; // [g][2 n_ang][n_cell_total][2]
    double *beta;  // [g][2 n_ang][n_cell_total]
};

static __global__ void corrections_kernel(const CorrArgs a)
{
    const int64_t per_g = (int64_t)a.n_planes * a.n_ang * a.n_cell_plane * 2;
    const int64_t total = per_g * a.g_count;
    const int n_cell_total = a.n_plane_total * a.n_cell_plane;
    const int nslot        = 2 * a.n_ang;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r      = i;
        const int dir  = (int)(r & 1);
        r >>= 1;
        const int ic   = (int)(r % a.n_cell_plane);
        r /= a.n_cell_plane;
        const int ang  = (int)(r % a.n_ang);
        r /= a.n_ang;
        const int ipl  = (int)(r % a.n_planes);
        const int grel = (int)(r / a.n_planes);
        const int g    = a.g_begin + grel;
        const int plane = a.planes[ipl];
        const int u     = a.plane_unique[plane];
        const int first_reg = a.plane_first_reg[plane];
        const int rb    = a.uniq_reg_begin[u];
        const int nreg_u = a.uniq_reg_begin[u + 1] - rb;
        const int32_t *cb = a.cell_fsr_begin + (size_t)u * (a.n_cell_plane + 1);
        const double *L   = a.geom_len + (size_t)rb * a.n_geom + (size_t)a.ang_geom[ang] * nreg_u;
        const double rs   = a.ang_rsintheta[ang];
        double vol = 0.0, sig = 0.0, norm = 0.0;
        for (int j = cb[ic]; j < cb[ic + 1]; j++) {
            const int rl   = a.cell_fsr[rb + j];
            const int R    = rl + first_reg;
            const size_t o = (size_t)R * a.GP + g;
            const double T = rs * L[rl];
            size_t dofs    = (size_t)R * nslot + ang * 2 + dir;
            if (a.gl == 1)
                dofs += (size_t)grel * a.n_reg * nslot;
            else
                dofs = dofs * a.GP + g;
            const double fv = T * a.qbar[o] + a.dsum[dofs] / a.xstr[o];
            vol += fv;
            sig += a.xstr_true[o] * fv;
            norm += T;
        }
        sig /= vol;
        vol /= norm;
        // upwind/downwind faces of the cell for this direction (correction_worker.cpp:47-66); E,N,W,S = 0,1,2,3
        const bool pos_x = (a.ang_ox[ang] > 0.0) != (dir == 1);
        const int s_xl = pos_x ? 2 : 0, s_xr = pos_x ? 0 : 2;
        const int s_yl = dir ? 1 : 3, s_yr = dir ? 3 : 1;
        const double area_x = a.ang_area_x[ang] / a.cell_dx[ic];
        const double area_y = a.ang_area_y[ang] / a.cell_dy[ic];
        auto ss = [&](int face) {
            const int surf = a.coarse_surf[4 * ic + face];
            size_t so      = (((size_t)plane * a.n_ang + ang) * a.n_surf_plane + surf) * 2 + dir;
            if (a.gl == 1)
                so += (size_t)grel * a.n_plane_total * a.n_ang * a.n_surf_plane * 2;
            else
                so = so * a.GP + g;
            return a.ssum[so];
        };
        const double psi_xl = ss(s_xl) * area_x, psi_xr = ss(s_xr) * area_x;
        const double psi_yl = ss(s_yl) * area_y, psi_yr = ss(s_yr) * area_y;
        const double ax = vol / (psi_xl + psi_xr);
        const double ay = vol / (psi_yl + psi_yr);
        const int icc   = ic + a.plane_cell_offset[plane];
        const double b  = sig / a.sn_xs[(size_t)(plane * a.n_cell_plane + ic) * a.GP + g];
        const int iang  = ang + dir * a.n_ang;
        const size_t oc = ((size_t)grel * nslot + iang) * n_cell_total + icc;
        a.alpha[2 * oc]     = ax;
        a.alpha[2 * oc + 1] = ay;
        a.beta[oc]          = b;
    }
}

} // namespace mocb200
