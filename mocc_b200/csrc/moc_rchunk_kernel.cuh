// moc_rchunk_kernel.cuh -- per-group production sweep kernel, second generation (sm_100a):
// REGISTER-RESIDENT CHUNKS IN STATICALLY PACKED BATCHES.
//
// Restates the same reference loops as moc_chunk_kernel.cuh
//   sweep1g<CurrentWorker>              src/sweepers/moc/moc_sweeper_kernel.inc.hpp:84-133
//   moc::Current::post_ray              src/sweepers/moc/moc_current_worker.hpp:202-264
//   cmdo::CurrentCorrections::post_ray  src/sweepers/cmdo/correction_worker.hpp:109-205
//   BoundaryCondition::update           src/core/boundary_condition.cpp:155-191
// on the attenuation cache, one energy group per work item.
//
// What the first chunk kernel spent its time on (profiles/r1): 214 thread instructions and 77 shared-memory
// wavefronts per segment position -- every (e, q) re-read from shared memory by compose and by both walks,
// rolled loops with address arithmetic, one 5-step scan of three 64-bit values per polar angle for every
// track however short, per-track descriptors / work counters / mailboxes. Here:
//   * a track is cut into CHUNKS of exactly LMAX slots (the last one padded with neutral slots: attenuation
//     1, q-bar 0, FSR id -1); chunks of many tracks are packed at set-up into BATCHES of NC chunks so that
//     no track straddles a batch (first-fit decreasing; a track longer than a batch becomes a chained unit).
//     Every batch is the same amount of work, whatever the track lengths: no lane idles on short tracks;
//   * a TEAM of NW warps sweeps a batch: P lanes per chunk, one per polar angle of the bundle. A lane loads
//     its chunk from shared memory ONCE (compose), keeps 1-e and q-bar in registers (fully unrolled, no address
//     arithmetic) and walks it forward and backward from registers;
//   * the flux entering every chunk comes from ONE scan over the lanes of the team. Track boundaries need no
//     segmented scan: the first chunk of a track carries the map x -> 0 x + (A psi_in + B), which annihilates
//     whatever precedes it (and likewise the last chunk for the backward direction);
//   * staging as before -- attenuations and FSR ids by TMA bulk copies, q-bar gathered by 8-byte cp.async with
//     the lanes on CONSECUTIVE slots, the tally reduced by one red.global.add.f64 per slot, again striped --
//     but the next batch's attenuations and q-bar are requested as soon as compose has emptied the buffers
//     (FSR ids triple-buffered), so they have the whole scan + walk + reduction to arrive.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "moc_chunk_kernel.cuh"

namespace mocb200 {

constexpr int kRcHead = 1; // first chunk of a track: the forward sweep starts here (incoming boundary flux)
constexpr int kRcTail = 2; // last chunk of a track: the backward sweep starts here
constexpr int kRcHeadCont = 4; // head of a later sub-block of a chained track: forward flux carried by the team
constexpr int kRcTailCont = 8; // tail of an earlier sub-block: backward flux from the team's scratch

// slots a team of NW warps sweeps at once, P lanes per chunk
__host__ __device__ constexpr int rc_chunks(int P, int NW)
{
    return 32 * NW / P;
}
__host__ __device__ constexpr int rc_slots(int P, int LMAX, int NW)
{
    return rc_chunks(P, NW) * LMAX;
}
// int32 words of one batch header: FSR ids of the slots + one int4 per lane of the team (what the plain sweep
// copies) + two int4 per chunk for the tallying sweeps
__host__ __device__ constexpr int rc_header_ints_plain(int P, int LMAX, int NW)
{
    return rc_slots(P, LMAX, NW) + 4 * 32 * NW;
}
__host__ __device__ constexpr int rc_header_ints(int P, int LMAX, int NW)
{
    return rc_header_ints_plain(P, LMAX, NW) + 8 * rc_chunks(P, NW);
}
// dynamic shared memory of one team: attenuations [NS][P], q-bar [NS], contributions [NS][P], per lane the angle
// weight and the incoming boundary flux of the staged item (doubles), batch headers [3] (triple-buffered); tally
// variants keep the crossing lists in global memory
__host__ __device__ constexpr size_t rc_team_bytes(int P, int LMAX, int NW)
{
    return (size_t)rc_slots(P, LMAX, NW) * (size_t)(2 * P + 1) * sizeof(double) + 2 * 32 * (size_t)NW * sizeof(double) +
           3 * (size_t)rc_header_ints(P, LMAX, NW) * sizeof(int32_t);
}

struct RcArgs {
    const int2 *units; // {first batch, batches}: swept by one team (batches > 1: one track longer than a batch)
    int32_t n_units;
    const int2 *pinfo; // per plane of the list: {macroplane, start of its FSRs in the regrouped q-bar / tally layout}
    int32_t n_planes;
    const int32_t *plane_first_reg; // original FSR numbering (TALLY 2: psi_diff sums per FSR)
    const int32_t *seg_fsr;         // padded segment arrays, original plane-local FSR ids (TALLY 2)
    const int2 *chunk_trk; // per chunk: {track (index into tracks), first position of the chunk in its track}
    const ChunkUnit *tracks;
    // Per batch one record of rc_header_ints() int32, fetched by ONE bulk copy two items ahead:
    //   [NS] plane-local FSR id of every slot IN THE REGROUPED NUMBERING (FSRs that rays visit together share
    //        32-byte sectors: build_fsr_groups in moc_api.cu); padding: a valid id with the sign bit set (q-bar is
    //        gathered from it -- the slot's 1 - e is 0 --, nothing is tallied)
    //   [32 NW] int4 per lane (chunk, polar angle): {flags | sweep angle << 8, incoming boundary slot, encoded
    //        outgoing slot, -}; head chunk = forward in / backward out, tail chunk = backward in / forward out
    //   [NC] 2 x int4 per chunk, copied by the tallying sweeps only: {first forward crossing at or after the
    //        chunk's first node, first backward crossing at or after its last node (indices into `cross`), segments
    //        of its track (0: empty chunk), its first position in the track}, {its first position in the padded
    //        segment arrays (original FSR ids), -, -, -}
    const int32_t *batch_hdr;
    const double *cache;     // attenuation cache of this list [plane][g][slot][P]
    int64_t n_slots;
    int32_t cache_groups, cache_g0;
    const double *wt_v_st; // [n_plane][n_ang]
    int32_t n_ang, bc_per_group;
    int32_t g_begin, g_count, GP, n_reg; // n_reg: FSRs of the original numbering (dsum); n_regp: stride of q / tally
    int32_t n_regp;
    const double *q; // group-major [g - g_begin][n_regp], regrouped numbering
    double *tally;   // same layout
    const double *bc_in;
    double *bc_out;
    int32_t interleave; // team numbering: 1: team * CTAs + CTA, 0: CTA * TEAMS + team
    double *scratch; // per team: backward flux entering each sub-block of a chained track, [sub-block][P]
    int32_t scratch_per_team;
    // coarse-mesh tallies of the last inner (TALLY 1: moc::Current, 2: cmdo::CurrentCorrections)
    const Cross *cross; // crossing lists with sentinels
    const double *cur_w, *flx_w; // [n_plane][n_ang][2]
    const int32_t *plane_surf_offset;
    double *current, *surface_flux; // [n_surf][GP]
    double *dsum, *ssum;
    int32_t n_surf_plane, n_plane_total;
};

// launch geometry: slots per chunk, warps per team, teams per CTA
struct RcConfig {
    int LMAX, NW, TEAMS;
};
typedef void (*RcFn)(const RcArgs);
// instantiated in moc_rc_p1.cu / moc_rc_p2.cu / moc_rc_p4.cu; nullptr: configuration not built
RcFn pick_rc_kernel_p1(int tally, const RcConfig &c);
RcFn pick_rc_kernel_p2(int tally, const RcConfig &c);
RcFn pick_rc_kernel_p4(int tally, const RcConfig &c);
struct RcPersistArgs;
typedef void (*RcPersistFn)(const RcPersistArgs);
RcPersistFn pick_rc_persist_kernel_p1(const RcConfig &c);
RcPersistFn pick_rc_persist_kernel_p2(const RcConfig &c);
RcPersistFn pick_rc_persist_kernel_p4(const RcConfig &c);

// One boundary phase of one inner sweep: every team walks its share of the work items (static round-robin).
// Called once per launch by sweep_rchunk_kernel and once per phase and inner by sweep_rchunk_persist_kernel.
//   s_bar, s_tot: static shared memory of the kernel (mbarriers initialised by the caller; their phase parities
//                 par_f / par_e live across calls)
//   bc_in/bc_out: boundary flux read / written by this phase
//   prestage:     what rc_prefetch_first has already requested for the team's first item (0: nothing,
//                 1: header and attenuations, 2: also the q-bar gather)
template <int P, int LMAX, int NW, int TEAMS, int TALLY, bool PREFETCH_ONLY = false>
__device__ __forceinline__ int rc_sweep_body(const RcArgs &a, uint64_t *s_bar, double (*s_tot)[NW][P][4], uint32_t &par_f,
                                              uint32_t &par_e, const double *bc_in, double *bc_out, int prestage)
{
    static_assert(P == 1 || P == 2 || P == 4, "one lane per polar angle: 1, 2 or 4 lanes per chunk");
    static_assert(LMAX % 2 == 1, "odd chunk length: conflict-free shared-memory strides");
    constexpr int T  = 32 * NW;                 // lanes of a team
    constexpr int NC = rc_chunks(P, NW);        // chunks per batch
    constexpr int NS = rc_slots(P, LMAX, NW);   // slots per batch
    constexpr int HS = rc_header_ints(P, LMAX, NW); // int32 words of a batch header
    extern __shared__ __align__(16) double s_dyn[];

    const int lane = threadIdx.x & 31;
    const int wid  = threadIdx.x >> 5;
    const int team = wid / NW, wl = wid - team * NW;
    const int tl   = wl * 32 + lane;     // lane within the team
    const int c    = tl / P, p = tl % P; // chunk within the batch, polar angle of the bundle
    const bool loader = tl == 0;
    char *wbase   = reinterpret_cast<char *>(s_dyn) + (size_t)team * rc_team_bytes(P, LMAX, NW);
    double *exb   = reinterpret_cast<double *>(wbase);
    double *qb    = exb + (size_t)NS * P;
    double *ab    = qb + NS;
    double *s_wt  = ab + (size_t)NS * P; // per lane: angle weight, incoming boundary flux of the staged item
    double *s_pin = s_wt + T;
    int32_t *fbuf = reinterpret_cast<int32_t *>(s_pin + T); // three batch-header buffers (FSR ids, lane descriptors)
    uint64_t *bar = &s_bar[4 * team];                                 // [0..2] header buffers, [3] attenuations
    auto team_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(T) : "memory"); };

    const int GP            = a.GP;
    const uint32_t per_unit = (uint32_t)a.n_planes * (uint32_t)a.g_count;
    const uint32_t total    = (uint32_t)a.n_units * per_unit;
    const uint32_t n_teams  = gridDim.x * TEAMS;
    // consecutive work items go to different SMs: the teams that get one item more are spread over all SMs
    const uint32_t team_global = a.interleave ? (uint32_t)team * gridDim.x + blockIdx.x : blockIdx.x * TEAMS + team;

    // work item -> (unit, plane of the list, group of the launch)
    struct Item {
        int batch, nb; // first batch, batches of the unit
        int ipl, grel, plane, first_reg;
    };
    auto decode = [&](uint32_t w) {
        Item it;
        uint32_t unit = w;
        it.ipl = 0, it.grel = 0;
        if (per_unit != 1u) {
            unit             = w / per_unit;
            const uint32_t r = w - unit * per_unit;
            it.ipl           = (int)(r / (uint32_t)a.g_count);
            it.grel          = (int)(r - (uint32_t)it.ipl * (uint32_t)a.g_count);
        }
        const int2 u = __ldg(a.units + unit);
        it.batch = u.x, it.nb = u.y;
        const int2 pi = __ldg(a.pinfo + it.ipl);
        it.plane = pi.x, it.first_reg = pi.y;
        return it;
    };
    // ---- staging ----
    auto issue_hdr = [&](int fi, int batch) {
        if (loader) {
            constexpr uint32_t bytes = (uint32_t)(TALLY ? HS : rc_header_ints_plain(P, LMAX, NW)) * 4u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar + fi, bytes);
            bulk_g2s(fbuf + fi * HS, a.batch_hdr + (size_t)batch * HS, bytes, bar + fi);
        }
    };
    auto issue_ex = [&](const Item &it, int batch) {
        if (loader) {
            constexpr uint32_t bytes = (uint32_t)NS * (uint32_t)P * 8u;
            const int g     = a.g_begin + it.grel;
            const char *src = reinterpret_cast<const char *>(
                a.cache + (((size_t)it.ipl * a.cache_groups + (g - a.cache_g0)) * a.n_slots + (size_t)batch * NS) * P);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar + 3, bytes);
            for (uint32_t off = 0; off < bytes; off += 16384u)
                bulk_g2s(reinterpret_cast<char *>(exb) + off, src + off, min(16384u, bytes - off), bar + 3);
        }
    };
    constexpr int NJ = (NS + T - 1) / T;
    // q-bar of the batch whose header sits in buffer fi, striped: lanes on consecutive slots. mt: the lane's
    // descriptor from the same header (read together with the FSR ids: one shared-memory latency for all of them)
    auto gather_q = [&](int fi, const Item &it, int4 &mt) {
        mbar_wait(bar + fi, (par_f >> fi) & 1u);
        par_f ^= 1u << fi;
        const double *qf = a.q + (size_t)it.grel * a.n_regp + it.first_reg;
        asm volatile("" : "+l"(qf)); // keep the base in a register pair: one IMAD.WIDE per address
        const int32_t *fb = fbuf + fi * HS;
        int f[NJ];
#pragma unroll
        for (int j = 0; j < NJ; j++) // all shared-memory reads first: one latency, not NJ
            f[j] = (NS % T == 0 || tl + j * T < NS) ? fb[tl + j * T] : 0;
        mt = *reinterpret_cast<const int4 *>(fb + NS + 4 * tl);
#pragma unroll
        for (int j = 0; j < NJ; j++)
            if (NS % T == 0 || tl + j * T < NS)
                cp_async_8(qb + tl + j * T, qf + (uint32_t)(f[j] & 0x7fffffff));
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto wait_staged = [&]() {
        mbar_wait(bar + 3, par_e);
        par_e ^= 1u;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        team_sync();
    };
    // Angle weight and incoming boundary flux of the lane for a staged item: asynchronous 8-byte copies into the
    // team's shared memory, complete with the q-bar gather (wait_staged), read by take_inputs. (Loading them into
    // registers one item ahead made the consumer wait for the NEXT item's loads as well: the loop's loads share a
    // scoreboard -- 11 % of the stall samples, profiles/r2/tuning.md.) The boundary values read here are never
    // written by the same phase.
    auto stage_inputs = [&](const Item &it, const int4 &mt) {
        cp_async_8(s_wt + tl, a.wt_v_st + it.plane * a.n_ang + (mt.x >> 8));
        if (mt.x & (kRcHead | kRcTail))
            cp_async_8(s_pin + tl, bc_in + ((size_t)it.plane * a.bc_per_group + mt.y) * GP + (a.g_begin + it.grel));
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto take_inputs = [&](const int4 &mt, double &wt_o, double &pin_o) { // behind wait_staged
        wt_o  = s_wt[tl];
        pin_o = (mt.x & (kRcHead | kRcTail)) ? s_pin[tl] : 0.0;
    };
    auto reduce_tally = [&](int fi, const Item &it) { // striped again: one red per slot
        double *tf = a.tally + (size_t)it.grel * a.n_regp + it.first_reg;
        asm volatile("" : "+l"(tf));
        const int32_t *fb = fbuf + fi * HS;
        int f[NJ];
        double v[NJ];
#pragma unroll
        for (int j = 0; j < NJ; j++) {
            const int i = tl + j * T;
            f[j] = -1, v[j] = 0.0;
            if (NS % T == 0 || i < NS) {
                f[j] = fb[i];
                if constexpr (P == 2) {
                    const double2 t = *reinterpret_cast<const double2 *>(ab + 2 * i);
                    v[j] = t.x + t.y;
                } else if constexpr (P == 4) {
                    const double2 t = *reinterpret_cast<const double2 *>(ab + 4 * i);
                    const double2 u = *reinterpret_cast<const double2 *>(ab + 4 * i + 2);
                    v[j] = (t.x + t.y) + (u.x + u.y);
                } else {
                    v[j] = ab[i];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NJ; j++) // predicated, explicit state space (a generic atomicAdd on the laundered pointer
                                     // would be an ATOM with a result; an `if` around it costs a branch per slot)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %2, 0;\n\t@p red.global.add.f64 [%0], %1;\n\t}" ::"l"(
                             tf + (uint32_t)(f[j] & 0x7fffffff)),
                         "d"(v[j]), "r"(f[j])
                         : "memory");
    };

    // ---- the lane's chunk: load once, keep 1 - e and q-bar in registers ----
    struct Maps {
        double Af, Bf, Ab, Bb;
    };
    auto compose = [&](double (&ome)[LMAX], double (&qv)[LMAX]) {
        const double *el = exb + (size_t)c * LMAX * P + p;
        const double *ql = qb + c * LMAX;
        double A = 1.0, Bf = 0.0, Bb = 0.0;
#pragma unroll
        for (int k = 0; k < LMAX; k++) {
            const double e = el[k * P];
            qv[k]  = ql[k];
            ome[k] = 1.0 - e;
            const double bq = qv[k] * ome[k];
            Bb = fma(A, bq, Bb); // M o m_k: the backward sweep applies the higher slot first
            Bf = fma(e, Bf, bq); // m_k o M
            A *= e;
        }
        return Maps{A, Bf, A, Bb};
    };
    // Scan over the chunks of the team. In: the forward / backward map of the lane's chunk (with the resets of
    // track heads / tails folded in). Out: the flux entering the chunk in both directions (valid for chunks that
    // are not heads / tails themselves); team_total: the team's total maps (chained tracks).
    // Two halves around ONE team barrier (the caller's: it also frees the staging buffers): scan_warp -- butterfly
    // over the lanes of the warp, the warp's total maps to shared memory; scan_team -- the other warps' totals.
    struct ScanState {
        double EfA, EfB, EbA, EbB, TAf, TF, TAb, TB;
    };
    auto scan_warp = [&](const Maps &m, ScanState &st) {
        double EfA = 1.0, EfB = 0.0, EbA = 1.0, EbB = 0.0;
        double TAf = m.Af, TF = m.Bf, TAb = m.Ab, TB = m.Bb;
#pragma unroll
        for (int s = P; s < 32; s <<= 1) {
            const bool upper = (lane & s) != 0; // the partner block holds the LOWER chunks
            const double oAf = __shfl_xor_sync(0xffffffffu, TAf, s);
            const double oF  = __shfl_xor_sync(0xffffffffu, TF, s);
            const double oAb = __shfl_xor_sync(0xffffffffu, TAb, s);
            const double oB  = __shfl_xor_sync(0xffffffffu, TB, s);
            if (upper) {
                EfB = fma(EfA, oF, EfB); // forward: the lower block first, then what I already have below me
                EfA *= oAf;
                TF = fma(TAf, oF, TF);   // merged block, forward: lower (partner) then upper (mine)
                TB = fma(oAb, TB, oB);   // backward: upper (mine) then lower (partner)
            } else {
                EbB = fma(EbA, oB, EbB); // backward: the upper block first, then what I already have above me
                EbA *= oAb;
                TF = fma(oAf, TF, oF);
                TB = fma(TAb, oB, TB);
            }
            TAf *= oAf;
            TAb *= oAb;
        }
        if (NW > 1 && lane < P) { // every lane of polar angle p holds the warp's total maps
            s_tot[team][wl][p][0] = TAf, s_tot[team][wl][p][1] = TF;
            s_tot[team][wl][p][2] = TAb, s_tot[team][wl][p][3] = TB;
        }
        st = ScanState{EfA, EfB, EbA, EbB, TAf, TF, TAb, TB};
    };
    auto scan_team = [&](const ScanState &st, double &psi_f, double &psi_b, Maps *team_total) {
        const double EfA = st.EfA, EfB = st.EfB, EbA = st.EbA, EbB = st.EbB;
        const double TAf = st.TAf, TF = st.TF, TAb = st.TAb, TB = st.TB;
        double cf = 0.0, cb = 0.0; // flux leaving the lower / higher warps (the first chunk of a batch is a head)
        if (NW > 1) {
#pragma unroll
            for (int w = 0; w < NW - 1; w++)
                if (w < wl)
                    cf = fma(s_tot[team][w][p][0], cf, s_tot[team][w][p][1]);
#pragma unroll
            for (int w = NW - 1; w > 0; w--)
                if (w > wl)
                    cb = fma(s_tot[team][w][p][2], cb, s_tot[team][w][p][3]);
            if (team_total) { // maps of the whole team (valid in every lane)
                double af = 1.0, bf = 0.0, ab_ = 1.0, bb = 0.0;
#pragma unroll
                for (int w = 0; w < NW; w++) {
                    bf = fma(s_tot[team][w][p][0], bf, s_tot[team][w][p][1]);
                    af *= s_tot[team][w][p][0];
                }
#pragma unroll
                for (int w = NW - 1; w >= 0; w--) {
                    bb = fma(s_tot[team][w][p][2], bb, s_tot[team][w][p][3]);
                    ab_ *= s_tot[team][w][p][2];
                }
                *team_total = Maps{af, bf, ab_, bb};
            }
        } else if (team_total) {
            *team_total = Maps{TAf, TF, TAb, TB};
        }
        psi_f = fma(EfA, cf, EfB);
        psi_b = fma(EbA, cb, EbB);
    };
    // both walks from registers (kernel:103-129); contributions of the lane's polar angle to ab[slot][p].
    // psi' = psi - (psi - q)(1 - e) as ONE dependent DADD + DFMA per slot (the reference's DADD, DMUL, DADD rounds
    // the product once more: 1e-16 relative); the tallied difference (psi - q)(1 - e) is computed off the chain
    auto walk = [&](const double (&ome)[LMAX], const double (&qv)[LMAX], double wt, double &psi_f, double &psi_b) {
        double s[LMAX];
#pragma unroll
        for (int k = 0; k < LMAX; k++) {
            const double t = psi_f - qv[k];
            psi_f = fma(-t, ome[k], psi_f);
            s[k]  = (t * ome[k]) * wt;
        }
        double *al = ab + (size_t)c * LMAX * P + p;
#pragma unroll
        for (int k = LMAX - 1; k >= 0; k--) {
            const double t = psi_b - qv[k];
            psi_b     = fma(-t, ome[k], psi_b);
            al[k * P] = fma(t * ome[k], wt, s[k]);
        }
    };
    // The walks of the last inner: as `walk`, plus moc::Current::post_ray (moc_current_worker.hpp:202-264) or
    // cmdo::CurrentCorrections::post_ray (correction_worker.hpp:109-205) at the coarse-surface crossings inside
    // the lane's chunk, for the lane's own polar angle (sweep angle `ang`).
    auto walk_tally = [&](const double (&ome)[LMAX], const double (&qv)[LMAX], double wt, double &psi_f, double &psi_b,
                          const Item &it, int batch, int fi_cur, int ang) {
        (void)batch;
        const int32_t *tdp = fbuf + fi_cur * HS + NS + 4 * T + 8 * c; // the chunk's tally descriptor
        const int4 td      = *reinterpret_cast<const int4 *>(tdp);
        const int nseg = td.z, k0 = td.w;
        const Cross *xfl = a.cross + td.x, *xbl = a.cross + td.y; // first crossings of the chunk, either direction
        const size_t wo  = ((size_t)it.plane * a.n_ang + ang) * 2;
        const double cw0 = a.cur_w[wo], cw1 = a.cur_w[wo + 1], fw0 = a.flx_w[wo], fw1 = a.flx_w[wo + 1];
        const int surf_off = a.plane_surf_offset[it.plane];
        const int g        = a.g_begin + it.grel;
        const int nslot    = 2 * a.n_ang;
        // psi_diff sums are kept per FSR of the ORIGINAL numbering (corrections_kernel reads them)
        const int32_t *fo = a.seg_fsr + (TALLY == 2 ? tdp[4] : 0);
        const int reg0    = TALLY == 2 ? a.plane_first_reg[it.plane] : 0;
        double *cur_g = a.current + (size_t)surf_off * GP + g, *sfl_g = a.surface_flux + (size_t)surf_off * GP + g;
        double *ssum_g = nullptr;
        if (TALLY == 2)
            ssum_g = a.ssum + (size_t)it.grel * a.n_plane_total * a.n_ang * a.n_surf_plane * 2 +
                     ((size_t)it.plane * a.n_ang + ang) * a.n_surf_plane * 2;
        auto tally_cross = [&](const Cross &x, double psi, int dir) {
            const bool ynorm = x.surf & 1;
            const uint32_t surf = (uint32_t)x.surf >> 1;
            const double cs = psi * (ynorm ? cw1 : cw0), fs = psi * (ynorm ? fw1 : fw0);
            // forward adds, backward subtracts (moc_current_worker.hpp:230-231); the corrections worker also
            // subtracts the backward SURFACE FLUX (correction_worker.hpp:136-137, 194-195)
            atomicAdd(cur_g + surf * (uint32_t)GP, dir ? -cs : cs);
            atomicAdd(sfl_g + surf * (uint32_t)GP, (dir && TALLY == 2) ? -fs : fs);
            if (TALLY == 2)
                atomicAdd(ssum_g + surf * 2u + dir, psi);
        };
        auto dsum_add = [&](int f, int dir, double d) {
            const size_t o = (size_t)it.grel * a.n_reg * nslot + (size_t)(f + reg0) * nslot + ang * 2 + dir;
            atomicAdd(&a.dsum[o], d);
        };
        // Positions [0, len) of the chunk are real. A forward crossing at track node n is tallied in front of position
        // n (by the chunk that holds it) or, n == nseg, behind the last position of the track: relative node
        // rf = n - k0 in [0, len), or len when the chunk ends the track. A backward crossing after nb walked segments
        // is tallied in front of position nseg - 1 - nb: rb = nseg - 1 - nb - k0 in [0, len), or -1 (nb == nseg:
        // behind position 0 of the track, k0 == 0). Padding slots leave the flux as it is, so the far-end checks may
        // sit on them. Crossings outside the chunk are switched off by a node no position compares equal to.
        const int len    = max(min(LMAX, nseg - k0), 0);
        const int lim_f  = (k0 + len == nseg) ? len : len - 1; // highest relative node this chunk tallies, forward
        int ci_f = 0, ci_b = 0;
        int rf = INT32_MAX, rb = INT32_MIN;
        // The crossing lists live in global memory. xf / xb: the crossing being waited for; f1, f2 / b1, b2: the two
        // behind it, requested before they are needed -- a crossing consumed inside the serial walk used to cost one
        // dependent L2 round trip each (the tallying inner ran 2.5x the plain one, profiles/r2/tuning.md). Reads run
        // up to three entries past a list's sentinel: the array is padded accordingly (build, moc_api.cu).
        Cross xf{INT32_MAX, 0}, xb{INT32_MAX, 0};
        Cross f1{INT32_MAX, 0}, f2{INT32_MAX, 0}, b1{INT32_MAX, 0}, b2{INT32_MAX, 0};
        auto set_f = [&]() {
            const int r = xf.node - k0;
            rf          = (xf.node != INT32_MAX && r <= lim_f) ? r : INT32_MAX;
        };
        auto set_b = [&]() {
            const int r = nseg - 1 - xb.node - k0; // -1 is the near end only for node == nseg (else: the chunk below)
            rb          = (xb.node != INT32_MAX && (r >= 0 || (r == -1 && xb.node == nseg))) ? r : INT32_MIN;
        };
        if (len > 0) {
            xf = xfl[0], xb = xbl[0];
            f1 = xfl[1], b1 = xbl[1];
            f2 = xfl[2], b2 = xbl[2];
            set_f();
            set_b();
        }
        auto next_f = [&]() {
            xf = f1, f1 = f2, f2 = xfl[ci_f + 3];
            ++ci_f;
            set_f();
        };
        auto next_b = [&]() {
            xb = b1, b1 = b2, b2 = xbl[ci_b + 3];
            ++ci_b;
            set_b();
        };
        // The walks themselves stay free of divergent code: a crossing visited inside the unrolled walk made the whole
        // warp execute the tally at nearly every step for the one or two lanes that had a crossing there (the tallying
        // inner ran 2.5x the plain one: 4.4 k of 6.9 k stall samples, profiles/r2/tuning.md). Instead every lane leaves
        // the flux at the node in front of each of its slots in its own part of the contribution buffer (only this lane
        // reads it back; the contributions are written last) and visits its crossings in a loop of its own afterwards.
        double s[LMAX];
        double *al = ab + (size_t)c * LMAX * P + p;
#pragma unroll
        for (int k = 0; k < LMAX; k++) {
            al[k * P]      = psi_f; // forward flux at the node in front of position k
            const double d = (psi_f - qv[k]) * ome[k];
            psi_f -= d;
            s[k] = d * wt;
            if (TALLY == 2 && k < len)
                dsum_add(fo[k], 0, d);
        }
        while (rf != INT32_MAX) { // rf == LMAX: the far end of the ray behind a full last chunk
            tally_cross(xf, rf == LMAX ? psi_f : al[rf * P], 0);
            next_f();
        }
#pragma unroll
        for (int k = LMAX - 1; k >= 0; k--) {
            al[k * P]      = psi_b; // backward flux at the node in front of position k (coming from above)
            const double d = (psi_b - qv[k]) * ome[k];
            psi_b -= d;
            s[k] = fma(d, wt, s[k]);
            if (TALLY == 2 && k < len)
                dsum_add(fo[k], 1, d);
        }
        while (rb != INT32_MIN) { // rb == -1: the near end of the ray
            tally_cross(xb, rb == -1 ? psi_b : al[rb * P], 1);
            next_b();
        }
#pragma unroll
        for (int k = 0; k < LMAX; k++)
            al[k * P] = s[k];
    };
    auto store_exit = [&](const Item &it, int enc, double v) {
        if (enc != INT32_MIN) {
            double *bc_out_pl = bc_out + (size_t)it.plane * a.bc_per_group * GP + (a.g_begin + it.grel);
            bc_out_pl[(size_t)(enc >= 0 ? enc : -(enc + 1)) * GP] = enc >= 0 ? v : 0.0;
        }
    };
    auto load_in = [&](const Item &it, int slot) {
        return bc_in[((size_t)it.plane * a.bc_per_group + slot) * GP + (a.g_begin + it.grel)];
    };

    // ================= pipeline over the work items of this team (static round-robin) =================
    // Invariant at the top of a single-batch item `cur` with `staged`: its header is in buffer fi, its attenuations
    // and q-bar are on their way, its lane descriptor, angle weight and incoming boundary flux are in registers (or
    // on their way); if the next item is a single batch too, its header has been requested into buffer (fi + 1) % 3.
    uint32_t w_cur = team_global;
    if (w_cur >= total)
        return 0;
    Item cur = decode(w_cur);
    uint32_t w_nxt = w_cur + n_teams;
    Item nxt{};
    if (w_nxt < total)
        nxt = decode(w_nxt);
    int fi = 0; // header buffer of `cur`
    if constexpr (PREFETCH_ONLY) {
        // Requests for the team's first item that do not depend on what the other CTAs are still writing: its
        // header and attenuations (prestage 1) and, when q-bar is final already, the q-bar gather (prestage 2).
        if (cur.nb != 1 || prestage == 0)
            return 0;
        team_sync(); // the team's last reduction has left the header buffers
        issue_hdr(fi, cur.batch);
        issue_ex(cur, cur.batch);
        if (w_nxt < total && nxt.nb == 1)
            issue_hdr((fi + 1) % 3, nxt.batch);
        if (prestage == 2) {
            int4 mt;
            gather_q(fi, cur, mt);
        }
        return prestage;
    }
    int4 meta = make_int4(0, 0, INT32_MIN, 0);
    double wt = 0.0, pin = 0.0;
    bool staged = false;
    double *sc  = a.scratch + (size_t)team_global * a.scratch_per_team;
    if (prestage != 0) { // rc_sweep_body<..., true> has requested the first item (cur.nb == 1)
        if (prestage == 1)
            gather_q(fi, cur, meta);
        else
            meta = *reinterpret_cast<const int4 *>(fbuf + fi * HS + NS + 4 * tl);
        stage_inputs(cur, meta);
        staged = true;
    }

    while (true) {
        const bool have_nxt = w_nxt < total;
        if (!staged && cur.nb == 1) { // (re)start the pipeline at `cur`
            issue_hdr(fi, cur.batch);
            issue_ex(cur, cur.batch);
            if (have_nxt && nxt.nb == 1)
                issue_hdr((fi + 1) % 3, nxt.batch);
            gather_q(fi, cur, meta);
            stage_inputs(cur, meta);
            staged = true;
        }
        if (cur.nb == 1) {
            // ---------------- one batch: the common case ----------------
            const bool head = meta.x & kRcHead, tail = meta.x & kRcTail;
            const bool pre  = have_nxt && nxt.nb == 1;
            // Static round-robin over equal batches. (A work counter claimed two items ahead was measured 5 % slower on
            // C5G7-2D: the teams run at a uniform pace, there is no imbalance to win back; profiles/r2/tuning.md.)
            const uint32_t w_nn = w_nxt + n_teams;
            Item nn{};
            if (have_nxt && w_nn < total) // two items ahead: its descriptor loads fly during compose
                nn = decode(w_nn);
            wait_staged();
            take_inputs(meta, wt, pin);
            double ome[LMAX], qv[LMAX];
            Maps m = compose(ome, qv);
            if (head)
                m.Bf = fma(m.Af, pin, m.Bf), m.Af = 0.0;
            if (tail)
                m.Bb = fma(m.Ab, pin, m.Bb), m.Ab = 0.0;
            ScanState st;
            scan_warp(m, st);
            team_sync(); // the warps' totals are visible; attenuation, q-bar and input buffers are free
            // the next item: attenuations, q-bar, angle weight and incoming flux (its header was requested one item
            // earlier); the item after it: header
            int4 meta_n = make_int4(0, 0, INT32_MIN, 0);
            if (pre) {
                issue_ex(nxt, nxt.batch);
                if (w_nn < total && nn.nb == 1)
                    issue_hdr((fi + 2) % 3, nn.batch);
                gather_q((fi + 1) % 3, nxt, meta_n);
                stage_inputs(nxt, meta_n);
            }
            double psi_f, psi_b;
            scan_team(st, psi_f, psi_b, nullptr);
            if (head)
                psi_f = pin;
            if (tail)
                psi_b = pin;
            if constexpr (TALLY == 0)
                walk(ome, qv, wt, psi_f, psi_b);
            else
                walk_tally(ome, qv, wt, psi_f, psi_b, cur, cur.batch, fi, meta.x >> 8);
            if (tail)
                store_exit(cur, meta.z, psi_f);
            if (head)
                store_exit(cur, meta.z, psi_b);
            team_sync(); // contributions complete
            reduce_tally(fi, cur);
            if (!have_nxt)
                break;
            if (pre) {
                fi = (fi + 1) % 3;
                meta = meta_n;
            } else {
                staged = false;
                team_sync(); // what follows reuses the buffers at once
            }
            w_cur = w_nxt, cur = nxt, w_nxt = w_nn, nxt = nn;
        } else {
            // ---------------- a track longer than a batch: sub-blocks chained by a carried flux ----------------
            // Not pipelined. Sub-block b holds chunks [b NC, (b + 1) NC) of the track; its first chunk is the head
            // of the track (b == 0) or continues the forward sweep with the flux the team carries, its last chunk is
            // the tail of the track (b == nb - 1) or continues the backward sweep with the flux pass A left in `sc`.
            const int nb = cur.nb;
            auto stage_block = [&](int b, int4 &mt) {
                team_sync(); // the previous block's reduction has left the buffers
                issue_hdr(fi, cur.batch + b);
                issue_ex(cur, cur.batch + b);
                gather_q(fi, cur, mt);
                wait_staged();
            };
            // pass A, highest sub-block first: the backward flux entering sub-block b - 1 from above
            for (int b = nb - 1; b >= 1; --b) {
                int4 mt;
                stage_block(b, mt);
                double ome[LMAX], qv[LMAX];
                Maps m = compose(ome, qv);
                const int flags = mt.x & 0xff;
                if (flags & kRcTail) // the tail of the track: incoming boundary flux
                    m.Bb = fma(m.Ab, load_in(cur, mt.y), m.Bb), m.Ab = 0.0;
                if (flags & kRcTailCont) // written by this loop's previous trip
                    m.Bb = fma(m.Ab, sc[b * P + p], m.Bb), m.Ab = 0.0;
                double pf, pb;
                Maps tot;
                ScanState st;
                scan_warp(m, st);
                team_sync();
                scan_team(st, pf, pb, &tot);
                if (tl < P) // the total backward map has A = 0 (a tail was folded in): B is the flux leaving below
                    sc[(b - 1) * P + p] = tot.Bb;
            }
            // pass B, lowest sub-block first: forward chain, both walks, tally
            double cfk = 0.0;
            for (int b = 0; b < nb; ++b) {
                int4 mt;
                stage_block(b, mt);
                const int flags  = mt.x & 0xff;
                const double wtb = __ldg(a.wt_v_st + cur.plane * a.n_ang + (mt.x >> 8));
                double ome[LMAX], qv[LMAX];
                Maps m = compose(ome, qv);
                const bool head = flags & kRcHead, tail = flags & kRcTail;
                const bool start_f = flags & (kRcHead | kRcHeadCont), start_b = flags & (kRcTail | kRcTailCont);
                double pin_f = 0.0, pin_b = 0.0;
                if (start_f)
                    pin_f = head ? load_in(cur, mt.y) : cfk;
                if (start_b)
                    pin_b = tail ? load_in(cur, mt.y) : sc[b * P + p];
                if (start_f)
                    m.Bf = fma(m.Af, pin_f, m.Bf), m.Af = 0.0;
                if (start_b)
                    m.Bb = fma(m.Ab, pin_b, m.Bb), m.Ab = 0.0;
                double psi_f, psi_b;
                Maps tot;
                ScanState st;
                scan_warp(m, st);
                team_sync();
                scan_team(st, psi_f, psi_b, &tot);
                if (start_f)
                    psi_f = pin_f;
                if (start_b)
                    psi_b = pin_b;
                if constexpr (TALLY == 0)
                    walk(ome, qv, wtb, psi_f, psi_b);
                else
                    walk_tally(ome, qv, wtb, psi_f, psi_b, cur, cur.batch + b, fi, mt.x >> 8);
                if (tail)
                    store_exit(cur, mt.z, psi_f);
                if (head)
                    store_exit(cur, mt.z, psi_b);
                cfk = tot.Bf; // forward flux leaving the sub-block (total A = 0: a head was folded in)
                team_sync();
                reduce_tally(fi, cur);
            }
            team_sync();
            if (!have_nxt)
                break;
            w_cur = w_nxt, cur = nxt;
            w_nxt = w_cur + n_teams;
            if (w_nxt < total)
                nxt = decode(w_nxt);
            staged = false;
        }
    }
    return 0;
}

template <int NW, int TEAMS> __device__ __forceinline__ void rc_init_barriers(uint64_t *s_bar)
{
    const int wid = threadIdx.x >> 5, team = wid / NW;
    if ((int)threadIdx.x == team * NW * 32) { // the loader lane of every team
        for (int i = 0; i < 4; i++)
            mbar_init(&s_bar[4 * team + i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

// One boundary phase per launch (Gauss-Seidel: octants 1/3, then 2/4; Jacobi: everything).
template <int P, int LMAX, int NW, int TEAMS, int TALLY>
__global__ void __launch_bounds__(32 * NW * TEAMS, 1) sweep_rchunk_kernel(const RcArgs a)
{
    __shared__ uint64_t s_bar[4 * TEAMS];
    __shared__ double s_tot[TEAMS][NW][P][4]; // per warp and polar angle: forward map (A, B), backward map (A, B) of the warp's chunks
    rc_init_barriers<NW, TEAMS>(s_bar);
    uint32_t par_f = 0u, par_e = 0u; // mbarrier phase parities (bit i of par_f: header buffer i)
    rc_sweep_body<P, LMAX, NW, TEAMS, TALLY>(a, s_bar, s_tot, par_f, par_e, a.bc_in, a.bc_out, 0);
}

// ---------------------------------------------------------------------------------------------------------------
// PERSISTENT VARIANT: all plain (NoCurrent) inner iterations of one sweep(group) call in ONE cooperative launch.
//
// What the per-phase launches cost on C5G7-2D (profiles/r2): the two sweep kernels of an inner run 49 + 48 us, the
// events around them measure 108.6 us (launch gaps), and inside each kernel the SMs are busy 80 % of the elapsed
// cycles (CTA launch, a cold pipeline -- header, then the q-bar gather that needs it, then compose --, and the
// tail). Here one CTA per SM stays resident for the whole call:
//   for every inner: phase 0 | grid barrier | phase 1 | grid barrier | flux + next q-bar (the arithmetic of
//   finalize_next_q_kernel, moc_sweep_kernel.cuh) | grid barrier
// and before it arrives at a barrier every team requests what its first batch behind the barrier needs and no
// other CTA is still writing: header + attenuations always, the q-bar gather too in front of phase 1.
// Grid barrier: one arrival counter in global memory, release / acquire at gpu scope (the acquire also drops the
// SM's L1 lines, which matters for q-bar and the boundary flux: both are rewritten by other SMs between phases).
struct RcFinalizeArgs {
    int32_t n_reg, GP, g_begin, g_count, reg_lo, reg_hi, n_regp;
    double *tally;      // group-major [g][n_regp], regrouped numbering
    const double *xstr; // [n_reg][GP] ...
    const double *vol;
    double *qbar, *flux;
    const double *src, *xs_self, *xstr_src;
    double *q_out; // group-major q-bar the sweep gathers from
    const int32_t *perm;
};

struct RcPersistArgs {
    RcArgs ph[2];       // per boundary phase (bc_in / bc_out of these are ignored)
    int32_t n_phases;   // 2: Gauss-Seidel boundary update, 1: Jacobi
    int32_t n_inner;    // inners swept by this launch
    int32_t finalize_last; // 0: the last inner's flux is left to finalize_flux_q_kernel (no next q-bar)
    int32_t prefetch;   // requests in front of the barriers: 0 none, 1 header + attenuations, 2 + q-bar gather
    double *bc[2];      // Gauss-Seidel: bc[0] read and written; Jacobi: read bc[i & 1], write bc[1 - (i & 1)]
    RcFinalizeArgs fin;
    unsigned int *barrier; // zeroed by the host before the launch
};

__device__ __forceinline__ void rc_grid_sync(unsigned int *counter, unsigned int &epoch)
{
    __syncthreads();
    epoch++;
    if (threadIdx.x == 0) {
        const unsigned int target = epoch * gridDim.x;
        unsigned int seen;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        do { // relaxed polls (an acquire load would drop the L1 on every trip), one acquire fence at the end
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

// flux = tally/(xstr vol) + 4 pi q-bar; q-bar' = (src + flux xs_self)/(4 pi xstr_src); tally = 0: the statements
// of finalize_next_q_kernel (same non-contracted arithmetic), grid-strided over the CTAs of this launch.
// with_next_q == false: the flux only (finalize_flux_q_kernel)
__device__ __forceinline__ void rc_finalize(const RcFinalizeArgs &f, bool with_next_q)
{

    const int nr    = f.reg_hi - f.reg_lo;
    const int64_t n = (int64_t)nr * f.g_count;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int gi    = (int)(i / nr);
        const int r     = f.reg_lo + (int)(i - (int64_t)gi * nr);
        const int g     = f.g_begin + gi;
        const size_t o  = (size_t)r * f.GP + g;
        const size_t oo = f.perm ? (size_t)gi * f.n_regp + f.perm[r] : (size_t)gi * f.n_reg + r;
        const double t  = __ldcg(f.tally + oo);
        const double fl = __dadd_rn(__ddiv_rn(t, __dmul_rn(f.xstr[o], f.vol[r])), __dmul_rn(f.qbar[o], kFPi));
        f.flux[o]       = fl;
        if (with_next_q) {
            const double r_fpi_tr = __ddiv_rn(1.0, __dmul_rn(f.xstr_src[o], kFPi));
            const double q        = __dmul_rn(__dadd_rn(f.src[o], __dmul_rn(fl, f.xs_self[o])), r_fpi_tr);
            f.qbar[o]   = q;
            f.q_out[oo] = q;
            f.tally[oo] = 0.0;
        }
    }
}

template <int P, int LMAX, int NW, int TEAMS>
__global__ void __launch_bounds__(32 * NW * TEAMS, 1) sweep_rchunk_persist_kernel(const __grid_constant__ RcPersistArgs pa)
{
    __shared__ uint64_t s_bar[4 * TEAMS];
    __shared__ double s_tot[TEAMS][NW][P][4];
    rc_init_barriers<NW, TEAMS>(s_bar);
    uint32_t par_f = 0u, par_e = 0u;
    unsigned int epoch = 0u;
    const bool jacobi = pa.n_phases == 1;
    int staged = 0; // what has been requested for the first item of the coming phase
    for (int inner = 0; inner < pa.n_inner; inner++) {
        const int flip       = jacobi ? (inner & 1) : 0;
        const double *bc_in  = pa.bc[flip];
        double *bc_out       = pa.bc[jacobi ? 1 - flip : 0];
        staged = rc_sweep_body<P, LMAX, NW, TEAMS, 0>(pa.ph[0], s_bar, s_tot, par_f, par_e, bc_in, bc_out, staged);
        if (!jacobi) {
            staged = rc_sweep_body<P, LMAX, NW, TEAMS, 0, true>(pa.ph[1], s_bar, s_tot, par_f, par_e, bc_in, bc_out,
                                                                 pa.prefetch);
            rc_grid_sync(pa.barrier, epoch); // outgoing boundary flux of phase 0 is what phase 1 starts from
            staged = rc_sweep_body<P, LMAX, NW, TEAMS, 0>(pa.ph[1], s_bar, s_tot, par_f, par_e, bc_in, bc_out, staged);
        }
        const bool more = inner + 1 < pa.n_inner;
        if (!more && !pa.finalize_last)
            break;
        rc_grid_sync(pa.barrier, epoch); // every contribution to the tally has arrived
        rc_finalize(pa.fin, true);
        if (more) // q-bar is being rewritten: header and attenuations only
            staged = rc_sweep_body<P, LMAX, NW, TEAMS, 0, true>(pa.ph[0], s_bar, s_tot, par_f, par_e, bc_in, bc_out,
                                                                 pa.prefetch ? 1 : 0);
        if (more)
            rc_grid_sync(pa.barrier, epoch);
    }
}

// Fills the attenuation cache of a packed list for groups [g_begin, g_begin + g_count): the same table
// lookup of -xstr*len/sin(theta) as exp_cache_kernel (exponential.hpp:69-79), one thread per slot; padding
// slots get the identity (1.0). Layout [plane][g][slot][P].
struct RcCacheArgs {
    const int2 *chunk_trk;
    const ChunkUnit *tracks;
    const int4 *len_begin; // per track and polar angle: start of that angle's own (padded) segment lengths
    const int32_t *planes;
    int32_t n_planes, lmax;
    int64_t n_slots;
    const double *seg_len;
    const int32_t *seg_fsr;
    const double *ang_rsintheta;
    const int32_t *plane_first_reg;
    const double *xstr; // [n_reg][GP]
    int32_t g_begin, g_count, cache_groups, cache_g0, GP;
    double *cache;
    const double *exp_table;
    int32_t exp_n;
    double exp_min, exp_max;
};

template <int P> __global__ void __launch_bounds__(512, 1) rc_cache_kernel(const RcCacheArgs a)
{
    extern __shared__ __align__(16) double s_tab[];
    for (int i = threadIdx.x; i < a.exp_n + 2; i += blockDim.x)
        s_tab[i] = a.exp_table[i];
    __syncthreads();
    const double space  = (a.exp_max - a.exp_min) / (double)a.exp_n;
    const double rspace = 1.0 / space;
    const double c0     = -a.exp_min * rspace;
    const double xmax   = (double)a.exp_n;
    const int64_t total = a.n_slots * a.n_planes;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ipl     = (int)(i / a.n_slots);
        const int64_t s   = i - (int64_t)ipl * a.n_slots;
        const int chunk   = (int)(s / a.lmax);
        const int k       = (int)(s - (int64_t)chunk * a.lmax);
        const int2 ct     = a.chunk_trk[chunk];
        bool valid        = ct.x >= 0;
        int reg           = 0;
        double len[P], nrs[P];
#pragma unroll
        for (int p = 0; p < P; p++)
            len[p] = 0.0, nrs[p] = 0.0;
        if (valid) {
            const ChunkUnit &u = a.tracks[ct.x];
            valid              = ct.y + k < u.nseg;
            if (valid) {
                const int plane = a.planes[ipl];
                reg             = a.seg_fsr[u.seg_begin + ct.y + k] + a.plane_first_reg[plane];
                const int4 lb4  = a.len_begin[ct.x];
                const int lb[4] = {lb4.x, lb4.y, lb4.z, lb4.w};
#pragma unroll
                for (int p = 0; p < P; p++) {
                    len[p] = a.seg_len[lb[p] + ct.y + k];
                    nrs[p] = -a.ang_rsintheta[u.ang[p]];
                }
            }
        }
        for (int gi = 0; gi < a.g_count; gi++) {
            const int g = a.g_begin + gi;
            double ex[P];
            if (valid) {
                const double xs = a.xstr[(size_t)reg * a.GP + g];
#pragma unroll
                for (int p = 0; p < P; p++)
                    ex[p] = exp_interp(s_tab, xs * len[p] * nrs[p], c0, rspace, xmax);
            } else {
#pragma unroll
                for (int p = 0; p < P; p++)
                    ex[p] = 1.0;
            }
            double *dst = a.cache + (((size_t)ipl * a.cache_groups + (g - a.cache_g0)) * a.n_slots + s) * P;
            if constexpr (P == 2) {
                *reinterpret_cast<double2 *>(dst) = make_double2(ex[0], ex[1]);
            } else if constexpr (P == 4) {
                *reinterpret_cast<double2 *>(dst)     = make_double2(ex[0], ex[1]);
                *reinterpret_cast<double2 *>(dst + 2) = make_double2(ex[2], ex[3]);
            } else {
                dst[0] = ex[0];
            }
        }
    }
}

} // namespace mocb200
