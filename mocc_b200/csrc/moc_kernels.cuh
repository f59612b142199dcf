// moc_kernels.cuh -- device code of the B200 MoC transport sweep (sm_100a).
//
// Restates, B200-first, what the reference does in
//   sweep1g<CurrentWorker>            src/sweepers/moc/moc_sweeper_kernel.inc.hpp:36-180
//   Exponential_Linear<N>::exp        src/core/exponential.hpp:69-79
//   moc::Current::post_ray            src/sweepers/moc/moc_current_worker.hpp:202-264
//   BoundaryCondition::update         src/core/boundary_condition.cpp:155-191
//   SourceIsotropic::self_scatter     src/core/source_isotropic.cpp:21-43
//
// Execution model. A work ITEM is one (track, direction) of one polar bundle: a
// ray geometry shared by up to MOCB200_MAX_POLAR polar angles. A thread owns one
// (item, energy group); the groups of an item sit in consecutive lanes so that the
// segment stream (length, FSR id) is a warp-broadcast load and the per-FSR data
// (xstr, q-bar, tally), stored [n_reg][GP] group-fastest, is one coalesced
// 8*G-byte access. The polar angles of the bundle are independent dependency
// chains inside the thread (ILP) and their tally contributions are summed in
// registers before the single red.global.add.f64 per (segment, group).
// Warps pull items dynamically (longest rays first) from a global counter; CTAs
// are persistent (grid = SMs x CTAs/SM) so the 80 KB exponential table is staged
// into shared memory once per CTA per launch.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace mocb200 {

constexpr int kMaxPolar = 4;
constexpr double kPi    = 3.1415926535897932; // src/core/constants.hpp:21
constexpr double kFPi   = 4.0 * kPi;

// One (track, direction) of a polar bundle. 32 bytes, read as two 16-byte loads.
struct __align__(16) Item {
    int32_t seg_first; // first segment in WALK order (fwd: begin, bwd: begin+nseg-1)
    int32_t nseg;
    int32_t in_slot;  // boundary slot the walk starts from (Ray::bc(0) fwd / bc(1) bwd)
    int32_t out_slot; // boundary slot the walk ends on
    int32_t bundle;   // index into the bundle table
    int32_t dir;      // 0 forward, 1 backward
    int32_t cross_begin; // first coarse-surface crossing of this (track, dir)
    int32_t ncross;
};

// Coarse-surface crossing in walk order (moc::Current::post_ray unrolled at setup time)
struct __align__(8) Cross {
    int32_t node; // number of segments walked when the surface is crossed
    int32_t surf; // (plane-local surface index << 1) | normal (0 = X, 1 = Y)
};

struct Bundle {
    int32_t np;               // polar angles in this bundle
    int32_t ang[kMaxPolar];   // sweep-angle indices (octants 1-2)
};

// Everything the sweep kernel needs, passed by value (fits the 4 KB param space)
struct SweepArgs {
    // work list
    const Item *items;
    int32_t n_items;
    uint32_t *counter; // dynamic work counter (zeroed before launch)
    const Bundle *bundles;
    const int32_t *planes; // macroplanes sharing this ray set
    int32_t n_planes;
    // geometry
    const double *seg_len;
    const int32_t *seg_fsr;
    const Cross *cross;
    // angle tables
    const double *ang_rsintheta; // [n_ang]
    const double *wt_v_st;       // [n_plane][n_ang]
    const double *cur_w;         // [n_plane][n_ang][2]
    const double *flx_w;         // [n_plane][n_ang][2]
    const int32_t *bc_offset;    // [2*n_ang]
    const int32_t *bc_size_x;    // [2*n_ang]
    const int32_t *bc_dst_off;   // [2*n_ang][2]
    const int32_t *bc_dst_kind;  // [2*n_ang][2]
    const int32_t *plane_first_reg;
    const int32_t *plane_surf_offset;
    int32_t n_ang;
    int32_t bc_per_group;
    // group data, [n_reg][GP]
    int32_t g_begin, g_count, GP;
    const double *xstr;
    const double *qbar;
    double *tally;
    // boundary flux [n_plane][bc_per_group][GP]
    const double *bc_in;
    double *bc_out;
    // coarse tallies [n_surf][GP]
    double *current;
    double *surface_flux;
    // exponential table
    const double *exp_table;
    int32_t exp_n;
    double exp_min, exp_max;
};

// Exponential_Linear<N>::exp with the table in shared memory. Same operation order as
// exponential.hpp:76-78; out-of-range arguments fall back to exp() (the reference also
// prints a line there, which is dropped).
__device__ __forceinline__ double exp_table_lookup(const double *__restrict__ tab, double v, double vmin,
                                                   double vmax, double space, double rspace)
{
    if (v < vmin || v > vmax)
        return exp(v);
    int i    = (int)((v - vmin) * rspace);
    double r = v - (space * i + vmin);
    double d0 = tab[i];
    double d1 = tab[i + 1];
    return d0 + (d1 - d0) * r * rspace;
}

template <int P, int TALLY>
static __global__ void __launch_bounds__(512, 2) sweep_kernel(const SweepArgs a)
{
    extern __shared__ double s_tab[];
    for (int i = threadIdx.x; i < a.exp_n + 2; i += blockDim.x)
        s_tab[i] = a.exp_table[i];
    __syncthreads();

    const int lane       = threadIdx.x & 31;
    const int GP         = a.GP;
    const int gc         = a.g_count;
    const double space   = (a.exp_max - a.exp_min) / (double)a.exp_n;
    const double rspace  = 1.0 / space;
    const uint32_t per_plane = (uint32_t)a.n_items * (uint32_t)gc; // threads of work per plane
    const uint32_t total     = per_plane * (uint32_t)a.n_planes;
    uint32_t *counter        = a.counter;

    const double *__restrict__ seg_len  = a.seg_len;
    const int32_t *__restrict__ seg_fsr = a.seg_fsr;
    const double *__restrict__ xstr     = a.xstr;
    const double *__restrict__ qbar     = a.qbar;

    for (;;) {
        uint32_t base = 0;
        if (lane == 0)
            base = atomicAdd(counter, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total)
            break;
        uint32_t w = base + lane;
        if (w >= total)
            continue;
        const uint32_t ipl  = w / per_plane;
        w -= ipl * per_plane;
        const int plane     = a.planes[ipl];
        const int first_reg = a.plane_first_reg[plane];
        const int item_id   = (int)(w / (uint32_t)gc);
        const int g         = a.g_begin + (int)(w - (uint32_t)item_id * (uint32_t)gc);

        const int4 i0 = reinterpret_cast<const int4 *>(a.items)[2 * item_id];
        const int4 i1 = reinterpret_cast<const int4 *>(a.items)[2 * item_id + 1];
        const int seg_first = i0.x, nseg = i0.y, in_slot = i0.z, out_slot = i0.w;
        const int bundle = i1.x, dir = i1.y;
        const int step = dir ? -1 : 1;

        double psi[P], rs[P], wt[P];
        int ang_io[P]; // boundary angle index of this direction
        const double *bc_in_pl = a.bc_in + (size_t)plane * a.bc_per_group * GP;
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int ang = a.bundles[bundle].ang[p];
            ang_io[p]     = ang + dir * a.n_ang;
            rs[p]         = a.ang_rsintheta[ang];
            wt[p]         = a.wt_v_st[plane * a.n_ang + ang];
            psi[p]        = bc_in_pl[(size_t)(a.bc_offset[ang_io[p]] + in_slot) * GP + g];
        }

        // coarse-surface crossings (TALLY == 1)
        int ci = 0, ncross = 0, next_node = -1;
        const Cross *cr = nullptr;
        double cw[P][2], fw[P][2];
        if (TALLY == 1) {
            cr     = a.cross + i1.z;
            ncross = i1.w;
            next_node = ncross > 0 ? cr[0].node : -1;
#pragma unroll
            for (int p = 0; p < P; p++) {
                const int ang = a.bundles[bundle].ang[p];
                const size_t o = ((size_t)plane * a.n_ang + ang) * 2;
                // forward adds, backward subtracts (moc_current_worker.hpp:230-231)
                cw[p][0] = dir ? -a.cur_w[o] : a.cur_w[o];
                cw[p][1] = dir ? -a.cur_w[o + 1] : a.cur_w[o + 1];
                fw[p][0] = a.flx_w[o];
                fw[p][1] = a.flx_w[o + 1];
            }
        }
        const int surf_off = (TALLY == 1) ? a.plane_surf_offset[plane] : 0;

        // software pipeline: (len, fsr) two segments ahead, (xstr, qbar) one ahead
        int s         = seg_first;
        double len_c  = seg_len[s];
        int reg_c     = seg_fsr[s] + first_reg;
        double len_n  = 0.0;
        int reg_n     = reg_c;
        if (nseg > 1) {
            len_n = seg_len[s + step];
            reg_n = seg_fsr[s + step] + first_reg;
        }
        double xs_c = xstr[(size_t)reg_c * GP + g];
        double q_c  = qbar[(size_t)reg_c * GP + g];

        for (int k = 0; k < nseg; k++) {
            // stage A: fetch segment k+2
            double len_nn = 0.0;
            int reg_nn    = reg_n;
            if (k + 2 < nseg) {
                len_nn = seg_len[s + 2 * step];
                reg_nn = seg_fsr[s + 2 * step] + first_reg;
            }
            // stage B: fetch region data of segment k+1
            const double xs_n = xstr[(size_t)reg_n * GP + g];
            const double q_n  = qbar[(size_t)reg_n * GP + g];

            if (TALLY == 1) {
                while (next_node == k) {
                    const int sn   = cr[ci].surf;
                    const int norm = sn & 1;
                    const size_t o = (size_t)((sn >> 1) + surf_off) * GP + g;
                    double c = 0.0, f = 0.0;
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        c += psi[p] * cw[p][norm];
                        f += psi[p] * fw[p][norm];
                    }
                    atomicAdd(&a.current[o], c);
                    atomicAdd(&a.surface_flux[o], f);
                    ci++;
                    next_node = ci < ncross ? cr[ci].node : -1;
                }
            }

            // stage C: attenuate through segment k
            const double t = -xs_c * len_c;
            double acc     = 0.0;
#pragma unroll
            for (int p = 0; p < P; p++) {
                const double e   = 1.0 - exp_table_lookup(s_tab, t * rs[p], a.exp_min, a.exp_max, space, rspace);
                const double dps = (psi[p] - q_c) * e;
                psi[p] -= dps;
                acc += dps * wt[p];
            }
            atomicAdd(&a.tally[(size_t)reg_c * GP + g], acc);

            len_c = len_n, reg_c = reg_n, xs_c = xs_n, q_c = q_n;
            len_n = len_nn, reg_n = reg_nn;
            s += step;
        }

        if (TALLY == 1) {
            while (ci < ncross) { // crossings at the far end of the ray (node == nseg)
                const int sn   = cr[ci].surf;
                const int norm = sn & 1;
                const size_t o = (size_t)((sn >> 1) + surf_off) * GP + g;
                double c = 0.0, f = 0.0;
#pragma unroll
                for (int p = 0; p < P; p++) {
                    c += psi[p] * cw[p][norm];
                    f += psi[p] * fw[p][norm];
                }
                atomicAdd(&a.current[o], c);
                atomicAdd(&a.surface_flux[o], f);
                ci++;
            }
        }

        // outgoing boundary flux goes straight to where BoundaryCondition::update would copy it
        double *bc_out_pl = a.bc_out + (size_t)plane * a.bc_per_group * GP;
#pragma unroll
        for (int p = 0; p < P; p++) {
            const int ao   = ang_io[p];
            const int sx   = a.bc_size_x[ao];
            const int face = out_slot >= sx ? 1 : 0;
            const int idx  = out_slot - (face ? sx : 0);
            const int kind = a.bc_dst_kind[2 * ao + face];
            if (kind != 2) {
                const size_t o = (size_t)(a.bc_dst_off[2 * ao + face] + idx) * GP + g;
                bc_out_pl[o]   = (kind == 1) ? psi[p] : 0.0;
            }
        }
    }
}

// q-bar = (src + flux*xs_self) * (1/(xstr_src*4pi)); also clears the sweep tally.
// Non-contracted arithmetic: bit-identical to source_isotropic.cpp:29-31.
static __global__ void self_scatter_kernel(int n_reg, int GP, int g_begin, int g_count, const double *__restrict__ src,
                                    const double *__restrict__ flux, const double *__restrict__ xs_self,
                                    const double *__restrict__ xstr_src, double *__restrict__ qbar,
                                    double *__restrict__ tally, int compute_q)
{
    const int64_t n = (int64_t)n_reg * g_count;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int r    = (int)(i / g_count);
        const int g    = g_begin + (int)(i - (int64_t)r * g_count);
        const size_t o = (size_t)r * GP + g;
        if (compute_q) {
            const double r_fpi_tr = __ddiv_rn(1.0, __dmul_rn(xstr_src[o], kFPi));
            qbar[o] = __dmul_rn(__dadd_rn(src[o], __dmul_rn(flux[o], xs_self[o])), r_fpi_tr);
        }
        tally[o] = 0.0;
    }
}

// flux = tally/(xstr*vol) + qbar*4pi   (kernel:165-173)
static __global__ void finalize_flux_kernel(int n_reg, int GP, int g_begin, int g_count, const double *__restrict__ tally,
                                     const double *__restrict__ xstr, const double *__restrict__ vol,
                                     const double *__restrict__ qbar, double *__restrict__ flux,
                                     const int32_t *__restrict__ reg_mask_begin, int reg_lo, int reg_hi)
{
    const int64_t n = (int64_t)(reg_hi - reg_lo) * g_count;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int r    = reg_lo + (int)(i / g_count);
        const int g    = g_begin + (int)(i % g_count);
        const size_t o = (size_t)r * GP + g;
        flux[o] = __dadd_rn(__ddiv_rn(tally[o], __dmul_rn(xstr[o], vol[r])), __dmul_rn(qbar[o], kFPi));
    }
}

// host column layout [g_count][n] <-> device layout [n][GP]
static __global__ void scatter_columns_kernel(int64_t n, int GP, int g_begin, int g_count, const double *__restrict__ cols,
                                       double *__restrict__ dst)
{
    const int64_t tot = n * g_count;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gl = i / n, r = i - gl * n; // coalesced read of the column
        dst[r * GP + g_begin + gl] = cols[i];
    }
}
static __global__ void gather_columns_kernel(int64_t n, int GP, int g_begin, int g_count, const double *__restrict__ src,
                                      double *__restrict__ cols)
{
    const int64_t tot = n * g_count;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gl = i / n, r = i - gl * n;
        cols[i] = src[r * GP + g_begin + gl];
    }
}

static __global__ void zero_groups_kernel(int64_t n, int GP, int g_begin, int g_count, double *__restrict__ dst)
{
    const int64_t tot = n * g_count;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / g_count, gl = i - r * g_count;
        dst[r * GP + g_begin + gl] = 0.0;
    }
}

} // namespace mocb200
