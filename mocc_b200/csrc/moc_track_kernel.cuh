// moc_track_kernel.cuh -- the production transport-sweep kernel (sm_100a).
//
// Restates, B200-first, the inner loops of
//   sweep1g<CurrentWorker>            src/sweepers/moc/moc_sweeper_kernel.inc.hpp:84-133
//   Exponential_Linear<N>::exp        src/core/exponential.hpp:69-79
//   moc::Current::post_ray            src/sweepers/moc/moc_current_worker.hpp:202-264
//   BoundaryCondition::update         src/core/boundary_condition.cpp:155-191
//
// Execution model: ONE WARP PER TRACK. A track is one ray geometry shared by the
// polar angles of a bundle; the warp sweeps it in BOTH directions for all P polar
// angles of the bundle, so that every segment costs ONE red.global.add.f64 per group
// (forward + backward + all polar contributions are summed in registers first).
//
// The attenuation along a ray is an affine map per segment,
//     psi_out = a psi_in + b,   a = exp_table(-tau),  b = qbar (1 - a),
// and affine maps compose associatively. The warp therefore does not walk the ray
// serially: a lane owns C consecutive segments of a block of (32/GL)*C segments,
// composes its C maps, a warp-shuffle scan over the lanes yields the angular flux
// entering every lane's chunk, and each lane then walks only its own C segments
// exactly like the reference loop (psi_diff = (psi - qbar) e; psi -= psi_diff;
// tally += psi_diff w). Blocks of one track are chained through a carried flux.
// Because the backward direction enters a block from the far side, the kernel makes
// two passes over the track's blocks: pass 1 (last block to first) chains the
// backward flux and stores the flux entering each block; pass 2 (first to last)
// chains the forward flux, re-evaluates the maps and produces all tallies.
//
// GL lanes of a warp hold GL energy groups of the same segments (GL = 1 for the
// reference's per-group sweep(group) calls, 8 for group-batched sweeps); per-FSR data
// is stored [n_reg][GP] group-fastest so that those lanes make one coalesced access.
// Segment lengths / FSR ids are read with 16-byte vector loads: the C = 4 segments of
// a lane are contiguous and the lanes of a warp cover one contiguous block.
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
    int32_t pad0, pad1, pad2;
};

struct TrackArgs {
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
    // angle tables (see SweepArgs)
    const double *ang_rsintheta;
    const double *wt_v_st;
    const double *cur_w;
    const double *flx_w;
    const int32_t *bc_offset;
    const int32_t *bc_size_x;
    const int32_t *bc_dst_off;
    const int32_t *bc_dst_kind;
    const int32_t *plane_first_reg;
    const int32_t *plane_surf_offset;
    int32_t n_ang;
    int32_t bc_per_group;
    // group data
    int32_t g_begin, g_count, GP, n_gsets;
    const double2 *xq; // [n_reg][GP] {xstr, qbar}
    double *tally;     // [n_reg][GP]
    const double *bc_in;
    double *bc_out;
    double *current;
    double *surface_flux;
    // per-warp scratch: flux entering each block in the backward direction
    double *scratch;
    int32_t scratch_per_warp; // doubles
    // exponential table
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
// knots makes an index flip at an interval boundary harmless).
__device__ __forceinline__ double exp_interp(const double *__restrict__ tab, double v, double c0, double rspace)
{
    const double x = fma(v, rspace, c0); // (v - vmin) * rspace
    if (x < 0.0)                         // v < vmin: the reference falls back to std::exp
        return exp(v);
    const double magic = 4503599627370496.0; // 2^52
    const double xi    = __dadd_rd(x, magic);
    const int i        = __double2loint(xi);
    const double frac  = x - (xi - magic);
    const double d0    = tab[i];
    const double d1    = tab[i + 1];
    return fma(d1 - d0, frac, d0);
}

constexpr int kTrackBlock = 512; // threads per CTA (16 warps), one persistent CTA per SM

template <int GL, int P, int C, int TALLY>
__global__ void __launch_bounds__(kTrackBlock, 1) sweep_track_kernel(const TrackArgs a)
{
    static_assert(C == 4, "vector loads below assume 4 segments per lane");
    constexpr int NCH  = 32 / GL; // chunk lanes per warp
    constexpr int SEGB = NCH * C; // segments per block

    extern __shared__ __align__(16) double s_tab[];
    __shared__ uint64_t s_bar;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)(a.exp_n + 2) * 8u;
        mbar_expect_tx(&s_bar, bytes);
        // chunks of <= 32 KB keep every bulk copy well inside hardware limits
        for (uint32_t off = 0; off < bytes; off += 32768u) {
            const uint32_t n = min(32768u, bytes - off);
            bulk_g2s(reinterpret_cast<char *>(s_tab) + off, reinterpret_cast<const char *>(a.exp_table) + off, n,
                     &s_bar);
        }
    }
    mbar_wait(&s_bar, 0);

    const int lane = threadIdx.x & 31;
    const int ch   = lane / GL;
    const int gl   = lane - ch * GL;
    const int GP   = a.GP;
    const double space  = (a.exp_max - a.exp_min) / (double)a.exp_n;
    const double rspace = 1.0 / space;
    const double c0     = -a.exp_min * rspace;

    const uint32_t per_unit = (uint32_t)a.n_planes * (uint32_t)a.n_gsets;
    const uint32_t total    = (uint32_t)a.n_units * per_unit;
    const int warp_global   = (blockIdx.x * (kTrackBlock / 32)) + (threadIdx.x >> 5);
    double *sc              = a.scratch + (size_t)warp_global * a.scratch_per_warp;

    const double *__restrict__ seg_len  = a.seg_len;
    const int32_t *__restrict__ seg_fsr = a.seg_fsr;
    const double2 *__restrict__ xq      = a.xq;

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
        const int bundle = u1.x;

        int g          = a.g_begin + gset * GL + gl;
        const bool gok = g < a.g_begin + a.g_count;
        if (!gok)
            g = a.g_begin; // idle group lane: computes on valid addresses, never writes

        double nrs[P], wt[P], cf[P], cb[P];
        int ang[P];
        const double *bc_in_pl = a.bc_in + (size_t)plane * a.bc_per_group * GP;
#pragma unroll
        for (int p = 0; p < P; p++) {
            ang[p] = a.bundles[bundle].ang[p];
            nrs[p] = -a.ang_rsintheta[ang[p]];
            wt[p]  = a.wt_v_st[plane * a.n_ang + ang[p]];
            cf[p]  = bc_in_pl[(size_t)(a.bc_offset[ang[p]] + bc0) * GP + g];
            cb[p]  = bc_in_pl[(size_t)(a.bc_offset[ang[p] + a.n_ang] + bc1) * GP + g];
        }
        double cw[P][2], fw[P][2];
        int surf_off = 0;
        if (TALLY == 1) {
            surf_off = a.plane_surf_offset[plane];
#pragma unroll
            for (int p = 0; p < P; p++) {
                const size_t o = ((size_t)plane * a.n_ang + ang[p]) * 2;
                cw[p][0] = a.cur_w[o], cw[p][1] = a.cur_w[o + 1];
                fw[p][0] = a.flx_w[o], fw[p][1] = a.flx_w[o + 1];
            }
        }

        const int nblk = (nseg + SEGB - 1) / SEGB;

        // ================= pass 1: backward flux entering each block =================
        for (int b = nblk - 1; b >= 0; --b) {
            const int k0 = b * SEGB + ch * C;
            double len[C];
            int reg[C];
            if (k0 < nseg) {
                const double2 l01 = *reinterpret_cast<const double2 *>(seg_len + seg_begin + k0);
                const double2 l23 = *reinterpret_cast<const double2 *>(seg_len + seg_begin + k0 + 2);
                const int4 f      = *reinterpret_cast<const int4 *>(seg_fsr + seg_begin + k0);
                len[0] = l01.x, len[1] = l01.y, len[2] = l23.x, len[3] = l23.y;
                reg[0] = f.x, reg[1] = f.y, reg[2] = f.z, reg[3] = f.w;
            } else {
#pragma unroll
                for (int c = 0; c < C; c++)
                    len[c] = 0.0, reg[c] = 0;
            }
            if (ch == 0 && gok) {
#pragma unroll
                for (int p = 0; p < P; p++)
                    sc[(b * P + p) * GL + gl] = cb[p];
            }
            double2 v[C];
#pragma unroll
            for (int c = 0; c < C; c++)
                v[c] = xq[(size_t)(reg[c] + first_reg) * GP + g];
            double A[P], B[P];
#pragma unroll
            for (int p = 0; p < P; p++)
                A[p] = 1.0, B[p] = 0.0;
#pragma unroll
            for (int c = 0; c < C; c++) {
                const bool valid = k0 + c < nseg;
                const double t   = v[c].x * len[c];
#pragma unroll
                for (int p = 0; p < P; p++) {
                    double ex = exp_interp(s_tab, t * nrs[p], c0, rspace);
                    ex        = valid ? ex : 1.0;
                    const double bq = v[c].y * (1.0 - ex);
                    // M o m_c: the backward sweep applies the higher segment first
                    B[p] = fma(A[p], bq, B[p]);
                    A[p] *= ex;
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
            double len[C];
            int reg[C];
            int2 xp = make_int2(0, 0);
            if (k0 < nseg) {
                const double2 l01 = *reinterpret_cast<const double2 *>(seg_len + seg_begin + k0);
                const double2 l23 = *reinterpret_cast<const double2 *>(seg_len + seg_begin + k0 + 2);
                const int4 f      = *reinterpret_cast<const int4 *>(seg_fsr + seg_begin + k0);
                len[0] = l01.x, len[1] = l01.y, len[2] = l23.x, len[3] = l23.y;
                reg[0] = f.x + first_reg, reg[1] = f.y + first_reg, reg[2] = f.z + first_reg, reg[3] = f.w + first_reg;
                if (TALLY == 1)
                    xp = a.xptr[(seg_begin + k0) >> 2];
            } else {
#pragma unroll
                for (int c = 0; c < C; c++)
                    len[c] = 0.0, reg[c] = first_reg;
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
            double2 v[C];
#pragma unroll
            for (int c = 0; c < C; c++)
                v[c] = xq[(size_t)reg[c] * GP + g];

            double e[P][C];
            double A[P], Bf[P], Bb[P];
#pragma unroll
            for (int p = 0; p < P; p++)
                A[p] = 1.0, Bf[p] = 0.0, Bb[p] = 0.0;
#pragma unroll
            for (int c = 0; c < C; c++) {
                const bool valid = k0 + c < nseg;
                const double t   = v[c].x * len[c];
#pragma unroll
                for (int p = 0; p < P; p++) {
                    double ex = exp_interp(s_tab, t * nrs[p], c0, rspace);
                    ex        = valid ? ex : 1.0;
                    e[p][c]   = 1.0 - ex;
                    const double bq = v[c].y * e[p][c];
                    Bb[p] = fma(A[p], bq, Bb[p]); // M o m_c
                    Bf[p] = fma(ex, Bf[p], bq);   // m_c o M
                    A[p] *= ex;
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
                    // forward flux leaving the block (carried to the next one)
                    cf[p] = __shfl_sync(0xffffffffu, out_f, (NCH - 1) * GL + gl);
                }
            }

            // ---- walk the lane's own C segments like the reference loop ----
            double acc[C];
            Cross xf, xb;
            int ci_f = xp.x, ci_b = xp.y;
            if (TALLY == 1) {
                xf = a.cross[ci_f];
                xb = a.cross[ci_b];
            }
#pragma unroll
            for (int c = 0; c < C; c++) {
                if (TALLY == 1 && gok) {
                    const int node = k0 + c; // forward flux at the node in front of segment k0+c
                    while (xf.node == node && node < nseg) {
                        const int norm = xf.surf & 1;
                        const size_t o = (size_t)((xf.surf >> 1) + surf_off) * GP + g;
                        double cs = 0.0, fs = 0.0;
#pragma unroll
                        for (int p = 0; p < P; p++) {
                            cs = fma(psi_f[p], cw[p][norm], cs);
                            fs = fma(psi_f[p], fw[p][norm], fs);
                        }
                        atomicAdd(&a.current[o], cs);
                        atomicAdd(&a.surface_flux[o], fs);
                        xf = a.cross[++ci_f];
                    }
                }
                double s = 0.0;
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const double d = (psi_f[p] - v[c].y) * e[p][c];
                    psi_f[p] -= d;
                    s = fma(d, wt[p], s);
                }
                acc[c] = s;
                if (TALLY == 1 && gok && k0 + c == nseg - 1) { // far end of the ray
                    while (xf.node == nseg) {
                        const int norm = xf.surf & 1;
                        const size_t o = (size_t)((xf.surf >> 1) + surf_off) * GP + g;
                        double cs = 0.0, fs = 0.0;
#pragma unroll
                        for (int p = 0; p < P; p++) {
                            cs = fma(psi_f[p], cw[p][norm], cs);
                            fs = fma(psi_f[p], fw[p][norm], fs);
                        }
                        atomicAdd(&a.current[o], cs);
                        atomicAdd(&a.surface_flux[o], fs);
                        xf = a.cross[++ci_f];
                    }
                }
            }
#pragma unroll
            for (int c = C - 1; c >= 0; c--) {
                const int k = k0 + c;
                if (TALLY == 1 && gok && k < nseg) {
                    const int nb = nseg - 1 - k; // segments walked by the backward sweep so far
                    while (xb.node == nb) {
                        const int norm = xb.surf & 1;
                        const size_t o = (size_t)((xb.surf >> 1) + surf_off) * GP + g;
                        double cs = 0.0, fs = 0.0;
#pragma unroll
                        for (int p = 0; p < P; p++) {
                            cs = fma(psi_b[p], cw[p][norm], cs);
                            fs = fma(psi_b[p], fw[p][norm], fs);
                        }
                        atomicAdd(&a.current[o], -cs); // backward subtracts (moc_current_worker.hpp:231)
                        atomicAdd(&a.surface_flux[o], fs);
                        xb = a.cross[++ci_b];
                    }
                }
                double s = acc[c];
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const double d = (psi_b[p] - v[c].y) * e[p][c];
                    psi_b[p] -= d;
                    s = fma(d, wt[p], s);
                }
                if (gok && k < nseg)
                    atomicAdd(&a.tally[(size_t)reg[c] * GP + g], s);
                if (TALLY == 1 && gok && k == 0) { // near end of the ray
                    while (xb.node == nseg) {
                        const int norm = xb.surf & 1;
                        const size_t o = (size_t)((xb.surf >> 1) + surf_off) * GP + g;
                        double cs = 0.0, fs = 0.0;
#pragma unroll
                        for (int p = 0; p < P; p++) {
                            cs = fma(psi_b[p], cw[p][norm], cs);
                            fs = fma(psi_b[p], fw[p][norm], fs);
                        }
                        atomicAdd(&a.current[o], -cs);
                        atomicAdd(&a.surface_flux[o], fs);
                        xb = a.cross[++ci_b];
                    }
                }
            }
            // backward flux leaving the ray: lane of the first chunk of the first block
            if (b == 0) {
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

// q-bar = (src + flux*xs_self) * (1/(xstr_src*4pi)) written next to xstr in the interleaved
// {xstr, qbar} array the track kernel gathers from; also clears the sweep tally.
__global__ void self_scatter_xq_kernel(int n_reg, int GP, int g_begin, int g_count, const double *__restrict__ src,
                                       const double *__restrict__ flux, const double *__restrict__ xs_self,
                                       const double *__restrict__ xstr_src, const double *__restrict__ xstr,
                                       const double *qbar_in, double2 *__restrict__ xq, double *qbar, double *__restrict__ tally, int compute_q)
{
    const int64_t n = (int64_t)n_reg * g_count;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int r    = (int)(i / g_count);
        const int g    = g_begin + (int)(i - (int64_t)r * g_count);
        const size_t o = (size_t)r * GP + g;
        double q;
        if (compute_q) {
            const double r_fpi_tr = __ddiv_rn(1.0, __dmul_rn(xstr_src[o], kFPi));
            q = __dmul_rn(__dadd_rn(src[o], __dmul_rn(flux[o], xs_self[o])), r_fpi_tr);
        } else {
            q = qbar_in[o];
        }
        qbar[o]  = q;
        xq[o]    = make_double2(xstr[o], q);
        tally[o] = 0.0;
    }
}

} // namespace mocb200
