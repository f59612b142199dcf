// moc_cached_kernel.cuh -- track kernel that STREAMS the per-segment attenuation instead
// of re-evaluating the exponential table (sm_100a; the production path when the cache fits).
//
// Same algorithm and arithmetic as moc_track_kernel.cuh (one warp per track, both
// directions, affine scan over the lanes, one red.global.add.f64 per segment and group).
// The difference is where exp_table(-xstr*len/sin(theta)) comes from. For a given group
// the transport cross section only changes when the host uploads a new one (never in plain
// MoC, once per outer with 2D3D transverse-leakage splitting), while a sweep(group) call
// runs n_inner inner iterations of two passes each. The table values are therefore
// evaluated ONCE per cross-section upload by exp_cache_kernel (same shared-memory table,
// same interpolation, bit-identical values) into an HBM-resident array, and the sweep
// reads 8 bytes per (segment, polar angle, group) -- a coalesced stream B200's HBM3e
// delivers faster than the SMs can redo the lookups (2 bank-conflicted LDS + ~8 FP64
// ops each). The sweep then needs neither segment lengths nor cross sections: it streams
// FSR ids (4 B/segment) and attenuations, gathers q-bar and scatters the tally.
//
// GL = 1 (the reference's per-group sweep(group) contract): q-bar and the tally are
// single doubles scattered over the FSRs, so the kernel makes those accesses with the
// lanes on 32 CONSECUTIVE segments ("striped": neighbouring segments lie in the same
// pin, hence in the same few 128-byte lines of the group-major q/tally arrays) and
// converts to the lane-owns-4-consecutive-segments ("blocked") arrangement the affine
// scan needs through a padded, conflict-free shared-memory transpose.
// GL = 8 (group-batched): the 8 group lanes of a segment already make one coalesced
// 64-byte access to the [n_reg][GP] arrays; no transpose.
#pragma once

#include "moc_track_kernel.cuh"

namespace mocb200 {

struct CachedArgs {
    const TrackUnit *units; // pad0 = position of the unit's first (padded) segment inside the list's cache
    int32_t n_units;
    uint32_t *counter;
    const Bundle *bundles;
    const int32_t *planes;
    int32_t n_planes;
    const int32_t *seg_fsr; // padded
    const int2 *xptr;
    const Cross *cross;
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
    int32_t g_begin, g_count, GP, n_gsets, n_reg;
    // GL = 1: group-major q/tally [g][n_reg];  GL = 8: [n_reg][GP]
    const double *q;
    double *tally;
    const double *bc_in;
    double *bc_out;
    double *current;
    double *surface_flux;
    double *scratch;
    int32_t scratch_per_warp;
    // attenuation cache of this list: GL = 1: [plane][g][pos][P]; GL = 8: [plane][pos][P][GP]
    const double *cache;
    int64_t list_pseg;    // padded segments of all units of the list
    int32_t cache_groups; // groups the cache holds (GL = 1 layout)
};

constexpr int kCachedBlock = 512;
constexpr int kTransposeDoubles = 144; // 128 + 2 per 16: conflict-free 16-byte blocked reads

__device__ __forceinline__ int tpos(int s)
{
    return s + 2 * (s >> 4);
}

template <int GL, int P, int TALLY>
__global__ void __launch_bounds__(kCachedBlock, 1) sweep_cached_kernel(const CachedArgs a)
{
    constexpr int C    = 4;
    constexpr int NCH  = 32 / GL;
    constexpr int SEGB = NCH * C;
    __shared__ __align__(16) double s_tr[GL == 1 ? (kCachedBlock / 32) * kTransposeDoubles : 2];

    const int lane = threadIdx.x & 31;
    const int ch   = lane / GL;
    const int gl   = lane - ch * GL;
    const int GP   = a.GP;
    double *tr     = s_tr + (GL == 1 ? (threadIdx.x >> 5) * kTransposeDoubles : 0);

    const uint32_t per_unit = (uint32_t)a.n_planes * (uint32_t)a.n_gsets;
    const uint32_t total    = (uint32_t)a.n_units * per_unit;
    const int warp_global   = (blockIdx.x * (kCachedBlock / 32)) + (threadIdx.x >> 5);
    double *sc              = a.scratch + (size_t)warp_global * a.scratch_per_warp;
    const int32_t *__restrict__ seg_fsr = a.seg_fsr;

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
            g = a.g_begin;
        const int grel = g - a.g_begin;

        // per-FSR arrays and attenuation stream of this (plane, group)
        const double *__restrict__ qv;
        double *__restrict__ tv;
        const double *__restrict__ ex_base;
        if (GL == 1) {
            qv      = a.q + (size_t)grel * a.n_reg;
            tv      = a.tally + (size_t)grel * a.n_reg;
            ex_base = a.cache + (((size_t)ipl * a.cache_groups + g) * a.list_pseg + cpos) * P;
        } else {
            qv      = a.q + g;
            tv      = a.tally + g;
            ex_base = a.cache + (((size_t)ipl * a.list_pseg + cpos) * P) * GP + g;
        }

        double wt[P], cf[P], cb[P];
        int ang[P];
        const double *bc_in_pl = a.bc_in + (size_t)plane * a.bc_per_group * GP;
#pragma unroll
        for (int p = 0; p < P; p++) {
            ang[p] = a.bundles[bundle].ang[p];
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

        // loads q-bar (blocked, q[c] for the lane's own 4 segments) and the attenuations of a block
        auto load_block = [&](int b, double (&q)[C], double (&ex)[P][C], int (&fs)[C]) {
            const int k0 = b * SEGB + ch * C;
            if (GL == 1) {
                // striped: instruction i touches 32 consecutive segments
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
#pragma unroll
                for (int c = 0; c < C; c++)
                    q[c] = qv[(size_t)fs[c] * GP];
            }
            if (k0 < npad) {
                if (GL == 1) {
                    const double2 *src = reinterpret_cast<const double2 *>(ex_base + (size_t)k0 * P);
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
                            ex[p][c] = ex_base[((size_t)(k0 + c) * P + p) * GP];
                }
            } else {
#pragma unroll
                for (int c = 0; c < C; c++)
#pragma unroll
                    for (int p = 0; p < P; p++)
                        ex[p][c] = 1.0;
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
                    B[p] = fma(A[p], bq, B[p]);
                    A[p] *= ex[p][c];
                }
            }
#pragma unroll
            for (int s = GL; s < 32; s <<= 1) {
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const double Ao = __shfl_xor_sync(0xffffffffu, A[p], s);
                    const double Bo = __shfl_xor_sync(0xffffffffu, B[p], s);
                    if (lane & s)
                        B[p] = fma(Ao, B[p], Bo);
                    else
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
            if (TALLY == 1 && k0 < nseg)
                xp = a.xptr[(seg_begin + k0) >> 2];
            double eb[P];
#pragma unroll
            for (int p = 0; p < P; p++) {
                double x = 0.0;
                if (ch == 0 && gok)
                    x = sc[(b * P + p) * GL + gl];
                eb[p] = __shfl_sync(0xffffffffu, x, gl);
            }
            double e[P][C];
            double A[P], Bf[P], Bb[P];
#pragma unroll
            for (int p = 0; p < P; p++)
                A[p] = 1.0, Bf[p] = 0.0, Bb[p] = 0.0;
#pragma unroll
            for (int c = 0; c < C; c++) {
#pragma unroll
                for (int p = 0; p < P; p++) {
                    e[p][c]         = 1.0 - ex[p][c];
                    const double bq = q[c] * e[p][c];
                    Bb[p] = fma(A[p], bq, Bb[p]);
                    Bf[p] = fma(ex[p][c], Bf[p], bq);
                    A[p] *= ex[p][c];
                }
            }
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
                        if (ch >= s) {
                            Bf[p] = fma(Af[p], Be, Bf[p]);
                            Af[p] *= Ae;
                        }
                        if (ch + s < NCH) {
                            Bb[p] = fma(Ab[p], Bh, Bb[p]);
                            Ab[p] *= Ah;
                        }
                    }
                }
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const double out_f = fma(Af[p], cf[p], Bf[p]);
                    const double out_b = fma(Ab[p], eb[p], Bb[p]);
                    const double in_f  = __shfl_up_sync(0xffffffffu, out_f, GL);
                    const double in_b  = __shfl_down_sync(0xffffffffu, out_b, GL);
                    psi_f[p] = ch == 0 ? cf[p] : in_f;
                    psi_b[p] = ch == NCH - 1 ? eb[p] : in_b;
                    cf[p]    = __shfl_sync(0xffffffffu, out_f, (NCH - 1) * GL + gl);
                }
            }

            double acc[C];
            Cross xf, xb;
            int ci_f = xp.x, ci_b = xp.y;
            if (TALLY == 1) {
                xf = a.cross[ci_f];
                xb = a.cross[ci_b];
            }
            auto tally_cross = [&](const Cross &x, const double (&psi)[P], double sign) {
                const int norm = x.surf & 1;
                const size_t o = (size_t)((x.surf >> 1) + surf_off) * GP + g;
                double cs = 0.0, fsum = 0.0;
#pragma unroll
                for (int p = 0; p < P; p++) {
                    cs   = fma(psi[p], cw[p][norm], cs);
                    fsum = fma(psi[p], fw[p][norm], fsum);
                }
                atomicAdd(&a.current[o], sign * cs);
                atomicAdd(&a.surface_flux[o], fsum);
            };
#pragma unroll
            for (int c = 0; c < C; c++) {
                if (TALLY == 1 && gok) {
                    const int node = k0 + c;
                    while (xf.node == node && node < nseg) {
                        tally_cross(xf, psi_f, 1.0);
                        xf = a.cross[++ci_f];
                    }
                }
                double s = 0.0;
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const double d = (psi_f[p] - q[c]) * e[p][c];
                    psi_f[p] -= d;
                    s = fma(d, wt[p], s);
                }
                acc[c] = s;
                if (TALLY == 1 && gok && k0 + c == nseg - 1) {
                    while (xf.node == nseg) {
                        tally_cross(xf, psi_f, 1.0);
                        xf = a.cross[++ci_f];
                    }
                }
            }
#pragma unroll
            for (int c = C - 1; c >= 0; c--) {
                const int k = k0 + c;
                if (TALLY == 1 && gok && k < nseg) {
                    const int nb = nseg - 1 - k;
                    while (xb.node == nb) {
                        tally_cross(xb, psi_b, -1.0); // backward subtracts (moc_current_worker.hpp:231)
                        xb = a.cross[++ci_b];
                    }
                }
                double s = acc[c];
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const double d = (psi_b[p] - q[c]) * e[p][c];
                    psi_b[p] -= d;
                    s = fma(d, wt[p], s);
                }
                acc[c] = s;
                if (TALLY == 1 && gok && k == 0) {
                    while (xb.node == nseg) {
                        tally_cross(xb, psi_b, -1.0);
                        xb = a.cross[++ci_b];
                    }
                }
            }
            // ---- scalar-flux tally: one reduction per segment and group ----
            if (GL == 1) {
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
                        atomicAdd(&tv[(size_t)fs[c] * GP], acc[c]);
            }
            if (b == 0) {
#pragma unroll
                for (int p = 0; p < P; p++)
                    cb[p] = psi_b[p];
            }
        }

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
    const double *ang_rsintheta;
    const int32_t *plane_first_reg;
    const double *xstr; // [n_reg][GP]
    int32_t g_begin, g_count, g_cache_begin, g_cache_count, GP, np, group_major;
    double *cache;
    int64_t list_pseg;
    const double *exp_table;
    int32_t exp_n;
    double exp_min, exp_max;
};

__global__ void __launch_bounds__(512, 1) exp_cache_kernel(const CacheArgs a)
{
    extern __shared__ __align__(16) double s_tab[];
    for (int i = threadIdx.x; i < a.exp_n + 2; i += blockDim.x)
        s_tab[i] = a.exp_table[i];
    __syncthreads();
    const double space  = (a.exp_max - a.exp_min) / (double)a.exp_n;
    const double rspace = 1.0 / space;
    const double c0     = -a.exp_min * rspace;
    const int lane      = threadIdx.x & 31;
    const int warps     = gridDim.x * (blockDim.x >> 5);
    const int64_t total = (int64_t)a.n_units * a.n_planes;
    const int P         = a.np;
    for (int64_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += warps) {
        const int unit_id   = (int)(w / a.n_planes);
        const int ipl       = (int)(w - (int64_t)unit_id * a.n_planes);
        const int plane     = a.planes[ipl];
        const int first_reg = a.plane_first_reg[plane];
        const TrackUnit u   = a.units[unit_id];
        const int npad      = (u.nseg + 3) & ~3;
        for (int k = lane; k < npad; k += 32) {
            const bool valid = k < u.nseg;
            const double len = a.seg_len[u.seg_begin + k];
            const int reg    = a.seg_fsr[u.seg_begin + k] + first_reg;
            for (int gi = 0; gi < a.g_count; gi++) {
                const int g    = a.g_begin + gi;
                const double t = a.xstr[(size_t)reg * a.GP + g] * len;
                for (int p = 0; p < P; p++) {
                    const double nrs = -a.ang_rsintheta[a.bundles[u.bundle].ang[p]];
                    const double ex  = valid ? exp_interp(s_tab, t * nrs, c0, rspace) : 1.0;
                    size_t o;
                    if (a.group_major)
                        o = (((size_t)ipl * a.g_cache_count + (g - a.g_cache_begin)) * a.list_pseg + u.pad0 + k) * P + p;
                    else
                        o = (((size_t)ipl * a.list_pseg + u.pad0 + k) * P + p) * a.GP + g;
                    a.cache[o] = ex;
                }
            }
        }
    }
}

// q-bar (self scatter) into the layout the cached kernel reads, tally reset.
// group_major: q/tally are [g - g_begin][n_reg], else [n_reg][GP].
__global__ void self_scatter_q_kernel(int n_reg, int GP, int g_begin, int g_count, const double *__restrict__ src,
                                      const double *__restrict__ flux, const double *__restrict__ xs_self,
                                      const double *__restrict__ xstr_src, double *qbar, double *__restrict__ q_out,
                                      double *__restrict__ tally_out, int group_major, int compute_q)
{
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
        const size_t oo = group_major ? (size_t)gi * n_reg + r : o;
        q_out[oo]       = q;
        tally_out[oo]   = 0.0;
    }
}

// flux = tally/(xstr*vol) + qbar*4pi   (kernel:165-173), tally in either layout
__global__ void finalize_flux_q_kernel(int n_reg, int GP, int g_begin, int g_count, const double *__restrict__ tally,
                                       const double *__restrict__ xstr, const double *__restrict__ vol,
                                       const double *__restrict__ qbar, double *__restrict__ flux, int reg_lo,
                                       int reg_hi, int group_major)
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
        const size_t oo = group_major ? (size_t)gi * n_reg + r : o;
        flux[o] = __dadd_rn(__ddiv_rn(tally[oo], __dmul_rn(xstr[o], vol[r])), __dmul_rn(qbar[o], kFPi));
    }
}

} // namespace mocb200
