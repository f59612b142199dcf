// moc_api.cu -- the C ABI of include/mocc_b200.h: host-side set-up (work lists,
// coarse-crossing lists, uploads) and kernel orchestration. See moc_kernels.cuh for
// the device code and the execution model.
#include "mocc_b200.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "moc_kernels.cuh"
#include "moc_sweep_kernel.cuh"
#include "moc_chunk_kernel.cuh"
#include "moc_rchunk_kernel.cuh"

using namespace mocb200;

namespace {

thread_local std::string g_create_error;

struct DeviceBuf {
    void *p      = nullptr;
    size_t bytes = 0;
};

// Register-chunk kernel (moc_rchunk_kernel.cuh): the tracks of a list cut into chunks of LMAX slots and packed into
// batches of NC chunks; a unit is what one team sweeps (one batch, or the consecutive batches of one long track)
struct RcList {
    int P = 0, LMAX = 0, NW = 0, TEAMS = 0, NC = 0, NS = 0;
    int64_t n_slots = 0; // n_batches * NS: positions of the list's attenuation cache per (plane, group)
    int32_t n_batches = 0, n_units = 0, max_nb = 1;
    int2 *d_units = nullptr, *d_chunk_trk = nullptr, *d_pinfo = nullptr; // d_pinfo: {macroplane, start in the regrouped layout}
    int32_t *d_batch_hdr = nullptr; // per batch: FSR ids of the slots + one int4 per lane (rc_header_ints words)
};

struct TrackList { // one launch of the track kernel: units of one (unique plane, boundary phase, polar count)
    int unique = 0, phase = 0, np = 0;
    TrackUnit *d_units = nullptr;
    int32_t n_units    = 0;
    int32_t *d_planes  = nullptr;
    int32_t n_planes   = 0;
    int64_t segs       = 0; // reference segments (polar copies counted) swept per group, all planes
    int64_t pseg       = 0; // padded segments of all units (cache positions)
    int32_t max_nseg   = 0; // longest track of the list
    ChunkUnit *d_cunits = nullptr; // self-contained descriptors of the same units (chunk kernel)
    int2 *d_pinfo       = nullptr; // per plane of the list: {macroplane, first FSR}
    int4 *d_len_begin   = nullptr; // per unit and polar angle: start of that angle's own segment lengths
    std::vector<int32_t> nseg_desc; // track lengths, longest first (staging-cap choice of the chunk kernel)
    double *d_cache    = nullptr; // attenuation cache of this list (CACHED kernel)
    RcList rc;                    // register-chunk kernel: packed batches
    int64_t cache_positions(bool rchunk) const { return rchunk ? rc.n_slots : pseg; }
};

struct WorkList { // one kernel launch: items of one (unique plane, boundary phase, polar count)
    int unique = 0, phase = 0, np = 0;
    Item *d_items    = nullptr;
    int32_t n_items  = 0;
    int32_t *d_planes = nullptr;
    int32_t n_planes = 0;
    int64_t segs     = 0; // segments walked per group by this list (all its planes)
};

} // namespace

struct mocb200_sweeper {
    int device = 0;
    mocb200_options opt{};
    // dims
    int G = 0, GP = 0, n_reg = 0, n_plane = 0, n_ang = 0, bcpg = 0, n_surf = 0, n_surf_plane = 0;
    int plane_begin = 0, plane_end = 0;
    int reg_lo = 0, reg_hi = 0; // FSR range of this handle's planes
    int exp_n = 0;
    double exp_min = 0, exp_max = 0;
    int sm_count = 148;
    // device memory
    std::vector<void *> allocs;
    int64_t device_bytes = 0;
    double *d_seg_len = nullptr;
    int32_t *d_seg_fsr = nullptr;
    Cross *d_cross     = nullptr;
    Bundle *d_bundles  = nullptr; // strict bundles (item kernel)
    Bundle *d_tbundles = nullptr; // bundles of the track lists (topological when the attenuation cache is used)
    double *d_rsin = nullptr, *d_wt = nullptr, *d_curw = nullptr, *d_flxw = nullptr;
    int32_t *d_bc_offset = nullptr, *d_bc_size_x = nullptr, *d_bc_dst_off = nullptr, *d_bc_dst_kind = nullptr;
    int32_t *d_plane_first_reg = nullptr, *d_plane_surf_offset = nullptr;
    double *d_vol = nullptr, *d_exp = nullptr;
    double *d_xstr = nullptr, *d_xstr_src = nullptr, *d_xs_self = nullptr, *d_src = nullptr, *d_flux = nullptr,
           *d_qbar = nullptr, *d_tally = nullptr;
    double *d_bc[2]   = {nullptr, nullptr};
    int bc_cur        = 0;
    double *d_current = nullptr, *d_surfflux = nullptr;
    double *d_stage   = nullptr; // staging for column <-> [n][GP] transposes
    size_t stage_elems = 0;
    double *h_stage    = nullptr; // pinned
    // fused per-sweep transfers: one pinned + one device buffer each way
    double *h_in = nullptr, *d_in = nullptr, *h_out = nullptr, *d_out = nullptr;
    size_t io_elems = 0;
    cudaEvent_t ev_in = nullptr; // the last fused upload has left h_in
    bool ev_in_pending = false;
    uint32_t *d_counters = nullptr;
    int n_counters       = 0;
    std::vector<WorkList> lists;
    // track kernel (production path)
    int kernel = 0;       // MOCB200_KERNEL_* actually used
    int cache_layout = 0; // 0 = no cache allocated, 1 = group-major (GL 1), 8 = group-fastest (GL 8)
    std::vector<bool> cache_valid; // per group
    int cache_slots = 0;           // groups the group-major cache holds at once (G when everything fits)
    int cache_g0 = 0, cache_gn = 0; // groups resident when cache_slots < G: [cache_g0, cache_g0 + cache_gn)
    double *d_qg = nullptr, *d_tg = nullptr; // group-major q-bar / tally [G][n_reg] ([G][n_regp], regrouped FSRs: RCHUNK)
    int32_t *d_fsr_perm = nullptr;           // RCHUNK: FSR -> position in the regrouped layout
    int n_regp = 0;
    // 2D3D correction factors
    bool have_corr = false;
    std::string corr_why; // why MOCB200_TALLY_CORRECTIONS is unavailable
    std::vector<bool> have_sn_xs;
    int n_cell_plane = 0, n_geom = 0;
    int32_t *d_all_planes = nullptr; // macroplanes of this handle
    int32_t *d_plane_unique = nullptr, *d_plane_cell_offset = nullptr, *d_uniq_reg_begin = nullptr;
    int32_t *d_cell_fsr_begin = nullptr, *d_cell_fsr = nullptr, *d_ang_geom = nullptr, *d_coarse_surf = nullptr;
    double *d_geom_len = nullptr, *d_area_x = nullptr, *d_area_y = nullptr, *d_ox = nullptr, *d_cell_dx = nullptr,
           *d_cell_dy = nullptr, *d_sn_xs = nullptr;
    double *d_dsum = nullptr, *d_ssum = nullptr, *d_alpha = nullptr, *d_beta = nullptr;
    size_t dsum_elems = 0, ssum_elems = 0, corr_groups = 0;
    int corr_g_begin = 0, corr_g_count = 0; // groups the alpha/beta buffers currently hold
    std::vector<TrackList> tlists;
    double *d_pseg_len = nullptr; // segment arrays with every track padded to a multiple of 4
    int32_t *d_pseg_fsr = nullptr;
    int2 *d_xptr = nullptr;
    Cross *d_xcross = nullptr;
    double2 *d_xq = nullptr;
    double *d_scratch = nullptr;
    int scratch_per_warp = 0;
    int rc_scratch_per_team = 0;
    int max_nseg = 0;
    int track_grid = 0;
    std::vector<bool> have_xs;
    // streams / events
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
    bool timing   = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
    std::vector<int> ev_inners; // inner sweeps an event pair brackets (persistent launches: several)
    size_t ev_used = 0;
    // persistent register-chunk sweep (all plain inners of a call in one cooperative launch)
    // device-side source construction (mocb200_set_source_xs)
    int n_mat = 0;
    int32_t *d_fsr_mat = nullptr, *d_scat_band = nullptr;
    double *d_mat_nf = nullptr, *d_mat_ch = nullptr, *d_mat_scat = nullptr, *d_fs = nullptr, *d_ext = nullptr;
    bool have_fs = false;
    int n_family = 1;            // angle families of the problem (angle_families)
    bool family_partial = false; // this handle sweeps a proper subset of them: flux needs the tallies of the others
    int rc_interleave = 1; // MOCB200_RC_INTERLEAVE=0|1 (A/B hook)
    int persist_mode = 0; // 0: one launch per boundary phase; 1: persistent; 2: + requests in front of the grid barriers
    unsigned int *d_gridbar = nullptr;
    bool persist_attr_set[5] = {false, false, false, false, false};
    // stats
    mocb200_stats stats{};
    std::string error;
};

namespace {

int fail(mocb200_sweeper *h, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h)
        h->error = buf;
    else
        g_create_error = buf;
    return code;
}

#define CUDA_TRY(h, call)                                                                                   \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess)                                                                              \
            return fail(h, MOCB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                          \
    } while (0)

template <class T> int dev_alloc(mocb200_sweeper *h, T **out, size_t count)
{
    void *p      = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CUDA_TRY(h, cudaMalloc(&p, bytes));
    h->allocs.push_back(p);
    h->device_bytes += (int64_t)bytes;
    *out = (T *)p;
    return MOCB200_OK;
}

template <class T> int dev_upload(mocb200_sweeper *h, T **out, const T *src, size_t count)
{
    int rc = dev_alloc(h, out, count);
    if (rc)
        return rc;
    if (count)
        CUDA_TRY(h, cudaMemcpy(*out, src, count * sizeof(T), cudaMemcpyHostToDevice));
    return MOCB200_OK;
}

template <class T> int dev_upload(mocb200_sweeper *h, T **out, const std::vector<T> &v)
{
    return dev_upload(h, out, v.data(), v.size());
}

int grid_for(int64_t n, int block, int sm_count)
{
    int64_t g = (n + block - 1) / block;
    return (int)std::max<int64_t>(1, std::min<int64_t>(g, (int64_t)sm_count * 8));
}

// Unroll moc::Current::post_ray (moc_current_worker.hpp:202-264) for one track into the
// list of coarse-surface crossings met in walk order, forward and backward.
void build_crossings(const mocb200_problem &p, int64_t t, int nseg, std::vector<Cross> &fw, std::vector<Cross> &bw)
{
    auto normal_of_local = [&](int s) { return s < p.nx * p.ny + (p.nx + 1) * p.ny ? 0 : 1; };
    int cell_fw = p.trk_cm_start[4 * t + 0], cell_bw = p.trk_cm_start[4 * t + 1];
    int surf_fw = p.trk_cm_start[4 * t + 2], surf_bw = p.trk_cm_start[4 * t + 3];
    int iseg_fw = 0, iseg_bw = nseg;
    fw.push_back(Cross{iseg_fw, (surf_fw << 1) | normal_of_local(surf_fw)});
    bw.push_back(Cross{nseg - iseg_bw, (surf_bw << 1) | normal_of_local(surf_bw)});
    for (int64_t k = p.trk_cm_begin[t]; k < p.trk_cm_begin[t + 1]; k++) {
        uint32_t c = p.cm_data[k];
        int s_fw = c & 0xF, s_bw = (c >> 4) & 0xF, n_fw = (c >> 8) & 0xFF, n_bw = (c >> 16) & 0xFF;
        if (s_fw != 7) { // Surface::INVALID
            iseg_fw += n_fw;
            int surf = p.coarse_surf[4 * cell_fw + s_fw];
            fw.push_back(Cross{iseg_fw, (surf << 1) | ((s_fw == 0 || s_fw == 2) ? 0 : 1)});
        }
        if (s_bw != 7) {
            iseg_bw -= n_bw;
            int surf = p.coarse_surf[4 * cell_bw + s_bw];
            bw.push_back(Cross{nseg - iseg_bw, (surf << 1) | ((s_bw == 0 || s_bw == 2) ? 0 : 1)});
        }
        if (s_fw < 4) {
            int nb  = p.coarse_nbr[4 * cell_fw + s_fw];
            cell_fw = nb < 0 ? cell_fw : nb;
        }
        if (s_bw < 4) {
            int nb  = p.coarse_nbr[4 * cell_bw + s_bw];
            cell_bw = nb < 0 ? cell_bw : nb;
        }
    }
}

// ---- register-chunk kernel: launch geometry and batch packing ----
RcFn pick_rc_kernel(int np, int tally, const RcConfig &c); // below
RcPersistFn pick_rc_persist_kernel(int np, const RcConfig &c);

constexpr int kRcSmemBudget = 232448 - 1280; // opt-in shared memory per CTA minus the kernels' static part (<= 1152 B)

bool rc_config_fits(int np, const RcConfig &c)
{
    return pick_rc_kernel(np, 0, c) != nullptr && (size_t)c.TEAMS * rc_team_bytes(np, c.LMAX, c.NW) <= (size_t)kRcSmemBudget;
}

// Launch geometry of a list. Measured on C5G7-2D (profiles/r2/tuning.md): the longest chunks that still leave the
// kernel without register spills win (13 slots: least scan work per slot), and three four-warp teams without spills
// beat four with. More lanes per chunk (polar angles) mean fewer chunks per batch: eight-warp teams then keep the
// long tracks inside one batch (a chained unit is not pipelined).
RcConfig rc_pick_config(int np, int max_nseg)
{
    static const char *force = getenv("MOCB200_RC_CFG"); // tuning hook: "LMAX,NW,TEAMS"
    if (force) {
        RcConfig c{0, 0, 0};
        if (sscanf(force, "%d,%d,%d", &c.LMAX, &c.NW, &c.TEAMS) == 3 && rc_config_fits(np, c))
            return c;
    }
    const RcConfig pref[] = {{13, 4, 3}, {13, 8, 2}, {13, 4, 2}, {11, 4, 3}, {11, 4, 2}, {7, 4, 4}};
    const RcConfig *first = nullptr;
    for (const RcConfig &c : pref) {
        if (!rc_config_fits(np, c))
            continue;
        if (!first)
            first = &c;
        if (max_nseg <= rc_slots(np, c.LMAX, c.NW))
            return c;
    }
    return first ? *first : RcConfig{7, 4, 4};
}

// Regroups the FSRs of one unique plane so that FSRs which rays visit within a few consecutive segments share
// 32-byte sectors (4 doubles) of the q-bar / tally arrays: the striped gather and reduction of the register-chunk
// kernel then touch ~14 instead of ~23 sectors per 32 slots on C5G7-2D (both are bound by sectors per request).
// Two rounds of heavy-edge matching on the co-occurrence graph (pairs, then pairs of pairs); groups of 4 first
// (sector-aligned), smaller groups after them. Returns new plane-local id per original plane-local id.
std::vector<int32_t> build_fsr_groups(const mocb200_problem &p, int u, int nreg)
{
    const int64_t t0 = p.geom_trk_begin[(size_t)u * p.n_geom], t1 = p.geom_trk_begin[(size_t)(u + 1) * p.n_geom];
    const int64_t nseg_u = p.trk_seg_begin[t1] - p.trk_seg_begin[t0];
    const int64_t stride = std::max<int64_t>(1, nseg_u * 3 / 24000000); // sample the tracks of very large planes
    struct Edge {
        int32_t a, b;
        float w;
    };
    auto collect = [&](auto &&emit) {
        for (int64_t t = t0; t < t1; t += stride) {
            const int64_t s0 = p.trk_seg_begin[t], n = p.trk_seg_begin[t + 1] - s0;
            for (int64_t k = 0; k < n; k++)
                for (int d = 1; d <= 3 && k + d < n; d++) {
                    const int32_t x = p.seg_fsr[s0 + k], y = p.seg_fsr[s0 + k + d];
                    if (x != y)
                        emit(std::min(x, y), std::max(x, y));
                }
        }
    };
    // weighted edge list via an open-addressing table (key = a * nreg + b)
    auto edges_of = [&](int n_nodes, auto &&for_each_pair) {
        size_t cap = 1u << 16;
        std::vector<uint64_t> keys(cap, UINT64_MAX);
        std::vector<float> wts(cap, 0.f);
        size_t used = 0;
        auto rehash = [&]() {
            std::vector<uint64_t> k2(cap * 4, UINT64_MAX);
            std::vector<float> w2(cap * 4, 0.f);
            for (size_t i = 0; i < cap; i++)
                if (keys[i] != UINT64_MAX) {
                    size_t j = (keys[i] * 0x9E3779B97F4A7C15ull) >> 20 & (cap * 4 - 1);
                    while (k2[j] != UINT64_MAX)
                        j = (j + 1) & (cap * 4 - 1);
                    k2[j] = keys[i], w2[j] = wts[i];
                }
            keys.swap(k2), wts.swap(w2), cap *= 4;
        };
        for_each_pair([&](int32_t a, int32_t b, float w) {
            const uint64_t key = (uint64_t)a * (uint64_t)n_nodes + (uint64_t)b;
            size_t j = (key * 0x9E3779B97F4A7C15ull) >> 20 & (cap - 1);
            while (keys[j] != UINT64_MAX && keys[j] != key)
                j = (j + 1) & (cap - 1);
            if (keys[j] == UINT64_MAX) {
                keys[j] = key;
                if (++used * 2 > cap) {
                    wts[j] += w;
                    rehash();
                    return;
                }
            }
            wts[j] += w;
        });
        std::vector<Edge> e;
        e.reserve(used);
        for (size_t i = 0; i < cap; i++)
            if (keys[i] != UINT64_MAX)
                e.push_back({(int32_t)(keys[i] / (uint64_t)n_nodes), (int32_t)(keys[i] % (uint64_t)n_nodes), wts[i]});
        return e;
    };
    auto match = [](int n_nodes, std::vector<Edge> &e, std::vector<int32_t> &super) { // returns number of supernodes
        std::stable_sort(e.begin(), e.end(), [](const Edge &x, const Edge &y) {
            return x.w != y.w ? x.w > y.w : (x.a != y.a ? x.a < y.a : x.b < y.b); // deterministic
        });
        std::vector<int32_t> mate(n_nodes, -1);
        for (const Edge &x : e)
            if (mate[x.a] < 0 && mate[x.b] < 0)
                mate[x.a] = x.b, mate[x.b] = x.a;
        super.assign(n_nodes, -1);
        int ns = 0;
        for (int r = 0; r < n_nodes; r++)
            if (super[r] < 0) {
                super[r] = ns;
                if (mate[r] >= 0)
                    super[mate[r]] = ns;
                ns++;
            }
        return ns;
    };
    std::vector<Edge> e1 = edges_of(nreg, [&](auto &&add) { collect([&](int32_t a, int32_t b) { add(a, b, 1.f); }); });
    std::vector<int32_t> pair_of, quad_of;
    const int n_pair = match(nreg, e1, pair_of);
    std::vector<Edge> e2 = edges_of(n_pair, [&](auto &&add) {
        for (const Edge &x : e1) {
            const int32_t a = pair_of[x.a], b = pair_of[x.b];
            if (a != b)
                add(std::min(a, b), std::max(a, b), x.w);
        }
    });
    const int n_quad = match(n_pair, e2, quad_of);
    // members of every group in original order; full groups first (sector-aligned), each class in order of its
    // lowest original id (keeps neighbouring pins in neighbouring cache lines)
    std::vector<std::vector<int32_t>> members(n_quad);
    for (int r = 0; r < nreg; r++)
        members[quad_of[pair_of[r]]].push_back(r);
    std::vector<int32_t> perm(nreg, 0);
    int32_t next = 0;
    for (int pass = 0; pass < 2; pass++)
        for (int g = 0; g < n_quad; g++)
            if ((members[g].size() == 4) == (pass == 0))
                for (int32_t r : members[g])
                    perm[r] = next++;
    return perm;
}

// Cuts the tracks of a list (cu: longest first) into chunks of LMAX slots and packs them into batches of NC
// chunks such that no track straddles a batch (best fit, decreasing); a track with more than NC chunks takes
// consecutive batches of its own (chained unit). Uploads the per-slot / per-lane tables the kernel reads.
int build_rc_list(mocb200_sweeper *h, TrackList &tl, const std::vector<ChunkUnit> &cu, const std::vector<int32_t> &pfsr,
                  const std::vector<int32_t> &lperm, const std::vector<int2> &pinfo_rc, const std::vector<Cross> &xcross)
{
    RcList &rc = tl.rc;
    const RcConfig cfg = rc_pick_config(tl.np, tl.max_nseg);
    rc.P = tl.np, rc.LMAX = cfg.LMAX, rc.NW = cfg.NW, rc.TEAMS = cfg.TEAMS;
    rc.NC = rc_chunks(rc.P, rc.NW), rc.NS = rc.NC * rc.LMAX;
    const int NC = rc.NC, L = rc.LMAX, P = rc.P;
    // chunks a batch may hold: NC; the chunk_cap test hook lowers it (the rest of every batch stays empty) so
    // that small test problems exercise the chained path
    int NCp = NC;
    if (h->opt.chunk_cap != 0)
        NCp = std::max(2, std::min(NC, std::abs(h->opt.chunk_cap) / L));
    struct Placed {
        int unit, batch, chunk0, nch;
    };
    std::vector<Placed> placed;
    std::vector<int2> units;
    int n_batches = 0;
    std::vector<int> nch(cu.size());
    for (size_t i = 0; i < cu.size(); i++)
        nch[i] = std::max(2, (cu[i].nseg + L - 1) / L); // >= 2: head and tail never share a chunk
    // chained units first (they are the longest tracks): consecutive batches of their own
    for (size_t i = 0; i < cu.size(); i++) {
        if (nch[i] <= NCp)
            continue;
        const int nb = (nch[i] + NCp - 1) / NCp;
        if (nch[i] - (nb - 1) * NCp < 2) // every sub-block needs a first and a last chunk of its own
            nch[i] = (nb - 1) * NCp + 2;
        placed.push_back({(int)i, n_batches, 0, nch[i]});
        units.push_back(make_int2(n_batches, nb));
        n_batches += nb;
        rc.max_nb = std::max(rc.max_nb, nb);
    }
    // the rest: best fit into the open batches (open_by_room[r]: batches with r free chunks)
    std::vector<std::vector<int>> open_by_room(NCp + 1);
    std::vector<int> used; // chunks used per normal batch (indexed from first_normal)
    const int first_normal = n_batches;
    for (size_t i = 0; i < cu.size(); i++) {
        if (nch[i] > NCp)
            continue;
        int r = nch[i];
        while (r <= NCp && open_by_room[r].empty())
            r++;
        int b;
        if (r > NCp) {
            b = (int)used.size();
            used.push_back(0);
            r = NCp;
        } else {
            b = open_by_room[r].back();
            open_by_room[r].pop_back();
        }
        placed.push_back({(int)i, first_normal + b, used[b], nch[i]});
        used[b] += nch[i];
        if (r - nch[i] > 0)
            open_by_room[r - nch[i]].push_back(b);
    }
    for (size_t b = 0; b < used.size(); b++)
        units.push_back(make_int2(first_normal + (int)b, 1));
    n_batches += (int)used.size();
    rc.n_batches = n_batches, rc.n_units = (int32_t)units.size();
    rc.n_slots   = (int64_t)n_batches * rc.NS;
    if (rc.n_slots >= (int64_t)INT32_MAX)
        return fail(h, MOCB200_ERR_INVALID, "register-chunk kernel: slot count exceeds 32-bit indexing");

    // padding slots: a valid FSR id (q-bar is gathered from it; the slot's 1 - e is 0) with the sign bit set
    std::vector<int32_t> slot_fsr((size_t)rc.n_slots, INT32_MIN);
    std::vector<int2> chunk_trk((size_t)n_batches * NC, make_int2(-1, 0));
    std::vector<int4> lane_meta((size_t)n_batches * NC * P, make_int4(0, 0, INT32_MIN, 0));
    for (const Placed &pl : placed) {
        const ChunkUnit &u = cu[pl.unit];
        const bool chained = pl.nch > NCp;
        for (int j = 0; j < pl.nch; j++) {
            // chained units: sub-block j / NCp, chunk j % NCp of it
            const int in_batch = chained ? j % NCp : pl.chunk0 + j;
            const int gchunk   = (pl.batch + (chained ? j / NCp : 0)) * NC + in_batch;
            const int k0       = j * L;
            chunk_trk[gchunk]  = make_int2(pl.unit, k0);
            for (int k = 0; k < L; k++)
                slot_fsr[(size_t)gchunk * L + k] = k0 + k < u.nseg ? lperm[pfsr[(size_t)u.seg_begin + k0 + k]]
                                                                   : (lperm[pfsr[(size_t)u.seg_begin + u.nseg - 1]] | INT32_MIN);
            int flags = 0;
            if (j == 0)
                flags |= kRcHead;
            else if (chained && in_batch == 0)
                flags |= kRcHeadCont;
            if (j == pl.nch - 1)
                flags |= kRcTail;
            else if (chained && in_batch == NCp - 1)
                flags |= kRcTailCont;
            for (int q = 0; q < P; q++) {
                int4 m = make_int4(flags | (u.ang[q] << 8), 0, INT32_MIN, 0);
                if (flags & kRcHead) // the forward sweep enters here, the backward sweep leaves
                    m.y = u.in_f[q], m.z = u.out_b[q];
                else if (flags & kRcTail)
                    m.y = u.in_b[q], m.z = u.out_f[q];
                lane_meta[(size_t)gchunk * P + q] = m;
            }
        }
    }
    int rc2;
    // one record per batch, fetched by one bulk copy: [NS] FSR ids, [32 NW] lane descriptors, [NC] tally descriptors
    const int HS = rc_header_ints(P, L, rc.NW), T = 32 * rc.NW;
    std::vector<int32_t> hdr((size_t)n_batches * HS, 0);
    const int32_t any_sentinel = (int32_t)xcross.size() - 1; // the lists end with sentinels
    auto lower_bound = [&](int32_t begin, int32_t n, int node) { // first index with xcross[i].node >= node
        int32_t a0 = 0, a1 = n;
        while (a0 < a1) {
            const int32_t m = (a0 + a1) >> 1;
            if (xcross[(size_t)begin + m].node < node)
                a0 = m + 1;
            else
                a1 = m;
        }
        return begin + a0;
    };
    for (int b = 0; b < n_batches; b++) {
        int32_t *rec = hdr.data() + (size_t)b * HS;
        std::copy(slot_fsr.begin() + (size_t)b * rc.NS, slot_fsr.begin() + (size_t)(b + 1) * rc.NS, rec);
        std::memcpy(rec + rc.NS, lane_meta.data() + (size_t)b * T, (size_t)T * sizeof(int4));
        int32_t *td = rec + rc.NS + 4 * T;
        for (int ch = 0; ch < NC; ch++, td += 8) {
            const int2 ct = chunk_trk[(size_t)b * NC + ch];
            td[0] = td[1] = any_sentinel;
            if (ct.x < 0)
                continue;
            const ChunkUnit &u = cu[ct.x];
            const int k0 = ct.y, kt_end = std::min(k0 + L, u.nseg);
            if (kt_end > k0) {
                td[0] = lower_bound(u.cross_begin, u.n_fw, k0);
                td[1] = lower_bound(u.cross_begin + u.n_fw + 1, u.n_bw, u.nseg - kt_end);
            }
            td[2] = u.nseg, td[3] = k0, td[4] = u.seg_begin + k0;
        }
    }
    if ((rc2 = dev_upload(h, &rc.d_units, units)) || (rc2 = dev_upload(h, &rc.d_chunk_trk, chunk_trk)) ||
        (rc2 = dev_upload(h, &rc.d_batch_hdr, hdr)) || (rc2 = dev_upload(h, &rc.d_pinfo, pinfo_rc)))
        return rc2;
    return MOCB200_OK;
}

int validate(const mocb200_problem *p)
{
    if (!p)
        return fail(nullptr, MOCB200_ERR_INVALID, "problem is NULL");
    if (p->n_group < 1 || p->n_reg < 1 || p->n_plane < 1 || p->n_unique < 1 || p->n_ang < 2 || p->n_geom < 1 ||
        p->bc_per_group < 1 || p->exp_n < 1 || p->n_ang != 2 * p->ndir_oct)
        return fail(nullptr, MOCB200_ERR_INVALID, "inconsistent problem dimensions");
    if (p->n_seg >= (int64_t)INT32_MAX || p->n_cm >= (int64_t)INT32_MAX / 2)
        return fail(nullptr, MOCB200_ERR_INVALID, "segment count exceeds 32-bit indexing (%lld)", (long long)p->n_seg);
    const void *req[] = {p->ang_geom, p->ang_rsintheta, p->wt_v_st, p->cur_wx, p->cur_wy, p->flx_wx, p->flx_wy,
                         p->bc_offset, p->bc_size_x, p->bc_size_y, p->bc_dst_off, p->bc_dst_kind,
                         p->geom_trk_begin, p->trk_seg_begin, p->trk_bc, p->trk_cm_begin, p->trk_cm_start,
                         p->seg_len, p->seg_fsr, p->cm_data, p->plane_unique, p->plane_first_reg,
                         p->plane_cell_offset, p->plane_surf_offset, p->coarse_surf, p->coarse_nbr, p->vol,
                         p->exp_table};
    for (const void *q : req)
        if (!q)
            return fail(nullptr, MOCB200_ERR_INVALID, "problem has a NULL array");
    for (int a = 0; a < p->n_ang; a++)
        if (p->ang_geom[a] < 0 || p->ang_geom[a] >= p->n_geom)
            return fail(nullptr, MOCB200_ERR_INVALID, "ang_geom out of range");
    for (int ip = 0; ip < p->n_plane; ip++)
        if (p->plane_unique[ip] < 0 || p->plane_unique[ip] >= p->n_unique)
            return fail(nullptr, MOCB200_ERR_INVALID, "plane_unique out of range");
    return MOCB200_OK;
}

// ANGLE FAMILIES (sharding one 2-D plane over ranks): sets of sweep angles closed under everything
// that couples angles inside a sweep -- the two directions of a track (a, a + n_ang), the polar copies the kernels
// bundle (same azimuth), and the boundary update (boundary_condition.cpp:155-191: the outgoing face of angle a feeds
// the incoming face of reflect(a, normal)). On a product quadrature a family is one azimuth of octant 1, its mirror
// image in octant 2, all their polar copies and their reverses: C5G7-2D has 8. A handle restricted to a family range
// sweeps its tracks exactly as the whole sweep does (Gauss-Seidel order included); what couples the ranks is the FSR
// tally (and the coarse tallies), summed over ranks between sweep and flux update.
int angle_families(const mocb200_problem &p, std::vector<int> &family)
{
    const int nab = 2 * p.n_ang;
    std::vector<int> parent(nab);
    for (int i = 0; i < nab; i++)
        parent[i] = i;
    auto find = [&](int x) {
        while (parent[x] != x)
            x = parent[x] = parent[parent[x]];
        return x;
    };
    auto unite = [&](int a, int b) {
        a = find(a), b = find(b);
        if (a != b)
            parent[std::max(a, b)] = std::min(a, b);
    };
    for (int a = 0; a < p.n_ang; a++)
        unite(a, a + p.n_ang);
    // polar copies: angles of one octant whose rays have the same boundary layout and the same azimuth (same X / Y
    // face sizes and the same ray spacing areas)
    for (int oct = 0; oct < 2; oct++)
        for (int a = oct * p.ndir_oct; a < (oct + 1) * p.ndir_oct; a++)
            for (int b = a + 1; b < (oct + 1) * p.ndir_oct; b++)
                if (p.bc_size_x[a] == p.bc_size_x[b] && p.bc_size_y[a] == p.bc_size_y[b] &&
                    p.ang_area_x[a] == p.ang_area_x[b] && p.ang_area_y[a] == p.ang_area_y[b])
                    unite(a, b);
    for (int ao = 0; ao < nab; ao++)
        for (int face = 0; face < 2; face++) {
            if (p.bc_dst_kind[2 * ao + face] == 2)
                continue;
            const int dst = p.bc_dst_off[2 * ao + face];
            for (int t = 0; t < nab; t++)
                if (dst >= p.bc_offset[t] && dst < p.bc_offset[t] + p.bc_size_x[t] + p.bc_size_y[t]) {
                    unite(ao, t);
                    break;
                }
        }
    family.assign(nab, -1);
    int n = 0;
    std::vector<int> id(nab, -1);
    for (int i = 0; i < nab; i++) {
        const int r = find(i);
        if (id[r] < 0)
            id[r] = n++;
        family[i] = id[r];
    }
    return n;
}

int build(mocb200_sweeper *h, const mocb200_problem &p)
{
    const mocb200_options &opt = h->opt;
    h->G = p.n_group;
    h->GP = p.n_group <= 2 ? p.n_group : ((p.n_group + 3) & ~3);
    h->n_reg = p.n_reg, h->n_plane = p.n_plane, h->n_ang = p.n_ang, h->bcpg = p.bc_per_group;
    h->n_surf = p.n_surf, h->n_surf_plane = p.n_surf_plane;
    h->exp_n = p.exp_n, h->exp_min = p.exp_min, h->exp_max = p.exp_max;
    h->plane_begin = opt.plane_begin, h->plane_end = opt.plane_end;
    if (h->plane_begin == 0 && h->plane_end == 0)
        h->plane_end = p.n_plane;
    if (h->plane_begin < 0 || h->plane_end > p.n_plane || h->plane_begin >= h->plane_end)
        return fail(h, MOCB200_ERR_INVALID, "bad plane range [%d, %d)", h->plane_begin, h->plane_end);
    if (opt.exp_mode != MOCB200_EXP_TABLE)
        return fail(h, MOCB200_ERR_INVALID, "exp_mode %d is not implemented (only MOCB200_EXP_TABLE)", opt.exp_mode);
    int max_polar = opt.max_polar <= 0 ? 2 : std::min<int>(opt.max_polar, kMaxPolar);

    // FSR range of this handle's planes (planes are stored contiguously, ascending)
    h->reg_lo = p.plane_first_reg[h->plane_begin];
    h->reg_hi = (h->plane_end < p.n_plane) ? p.plane_first_reg[h->plane_end] : p.n_reg;

    // ---- kernel selection: the attenuation cache needs 8 bytes per (padded segment, polar angle, group, plane) ----
    h->kernel = opt.kernel;
    if (h->kernel < MOCB200_KERNEL_AUTO || h->kernel > MOCB200_KERNEL_RCHUNK)
        return fail(h, MOCB200_ERR_INVALID, "unknown kernel selection %d", h->kernel);
    if (h->kernel == MOCB200_KERNEL_AUTO || h->kernel == MOCB200_KERNEL_CACHED || h->kernel == MOCB200_KERNEL_CHUNK ||
        h->kernel == MOCB200_KERNEL_RCHUNK) {
        int64_t bytes = 0;
        for (int u = 0; u < p.n_unique; u++) {
            int64_t n_planes_u = 0;
            for (int ip = h->plane_begin; ip < h->plane_end; ip++)
                n_planes_u += p.plane_unique[ip] == u;
            for (int a = 0; a < p.n_ang && n_planes_u; a++) {
                const size_t gi = (size_t)u * p.n_geom + p.ang_geom[a];
                for (int64_t t = p.geom_trk_begin[gi]; t < p.geom_trk_begin[gi + 1]; t++)
                    bytes += ((p.trk_seg_begin[t + 1] - p.trk_seg_begin[t] + 3) & ~(int64_t)3) * n_planes_u * h->GP * 8;
            }
        }
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(h, cudaMemGetInfo(&free_b, &total_b));
        // all groups if they fit, else as many as fit (the cache is then rebuilt for the groups of every sweep call)
        int slots = h->G;
        while (slots > 1 && (double)bytes * slots / h->G >= 0.7 * (double)free_b)
            slots--;
        if (opt.cache_groups > 0)
            slots = std::min<int>(slots, opt.cache_groups);
        const bool fits = (double)bytes * slots / h->G < 0.7 * (double)free_b;
        h->cache_slots  = slots;
        if (!fits && h->kernel != MOCB200_KERNEL_AUTO)
            return fail(h, MOCB200_ERR_INVALID, "attenuation cache (%lld MiB per group) does not fit in device memory",
                        (long long)((bytes / h->G) >> 20));
        h->kernel = !fits ? MOCB200_KERNEL_TRACK : (h->kernel == MOCB200_KERNEL_AUTO ? MOCB200_KERNEL_RCHUNK : h->kernel);
    }
    const bool cached_build = h->kernel == MOCB200_KERNEL_CACHED || h->kernel == MOCB200_KERNEL_CHUNK ||
                              h->kernel == MOCB200_KERNEL_RCHUNK;

    // ---- topology classes of the geometry classes ----
    // The reference keeps one ray set per (azimuth, polar) angle; polar copies of an azimuth visit the same
    // FSRs, boundary slots and coarse surfaces but may differ in the last bits of their segment lengths
    // (per-angle volume correction, ray_data.cpp). Kernels that read the attenuation CACHE never read lengths
    // (the cache is built per polar angle from that angle's own lengths), so for them polar copies are
    // bundled whenever the topology matches; kernels that read lengths need bit-identical rays.
    std::vector<int> topo_of(p.n_geom);
    for (int g = 0; g < p.n_geom; g++) {
        topo_of[g] = g;
        for (int g2 = 0; g2 < g && cached_build && topo_of[g] == g; g2++) {
            if (topo_of[g2] != g2)
                continue;
            bool same = true;
            for (int u = 0; u < p.n_unique && same; u++) {
                const int64_t a0 = p.geom_trk_begin[(size_t)u * p.n_geom + g], a1 = p.geom_trk_begin[(size_t)u * p.n_geom + g + 1];
                const int64_t b0 = p.geom_trk_begin[(size_t)u * p.n_geom + g2], b1 = p.geom_trk_begin[(size_t)u * p.n_geom + g2 + 1];
                same = (a1 - a0) == (b1 - b0);
                for (int64_t i = 0; i < a1 - a0 && same; i++) {
                    const int64_t ta = a0 + i, tb = b0 + i;
                    const int64_t na = p.trk_seg_begin[ta + 1] - p.trk_seg_begin[ta];
                    const int64_t nc = p.trk_cm_begin[ta + 1] - p.trk_cm_begin[ta];
                    same = na == p.trk_seg_begin[tb + 1] - p.trk_seg_begin[tb] &&
                           nc == p.trk_cm_begin[tb + 1] - p.trk_cm_begin[tb] &&
                           std::equal(p.trk_bc + 2 * ta, p.trk_bc + 2 * ta + 2, p.trk_bc + 2 * tb) &&
                           std::equal(p.trk_cm_start + 4 * ta, p.trk_cm_start + 4 * ta + 4, p.trk_cm_start + 4 * tb) &&
                           std::equal(p.seg_fsr + p.trk_seg_begin[ta], p.seg_fsr + p.trk_seg_begin[ta] + na,
                                      p.seg_fsr + p.trk_seg_begin[tb]) &&
                           std::equal(p.cm_data + p.trk_cm_begin[ta], p.cm_data + p.trk_cm_begin[ta] + nc,
                                      p.cm_data + p.trk_cm_begin[tb]);
                }
            }
            if (same)
                topo_of[g] = g2;
        }
    }

    // ---- polar bundles: angles of one octant that share a geometry (strict: bit-identical rays, for the
    //      kernels that read lengths; topological: for the kernels on the attenuation cache) ----
    std::vector<Bundle> bundles, tbundles;
    std::vector<int> bundle_phase, bundle_geom, tbundle_phase, tbundle_geom;
    auto make_bundles = [&](bool topo, std::vector<Bundle> &out, std::vector<int> &out_phase, std::vector<int> &out_geom) {
        for (int oct = 0; oct < 2; oct++) {
            for (int geom = 0; geom < p.n_geom; geom++) {
                std::vector<int> angs;
                for (int a = oct * p.ndir_oct; a < (oct + 1) * p.ndir_oct; a++)
                    if ((topo ? topo_of[p.ang_geom[a]] : p.ang_geom[a]) == geom)
                        angs.push_back(a);
                if (angs.empty())
                    continue;
                int nchunk = ((int)angs.size() + max_polar - 1) / max_polar;
                for (int c = 0; c < nchunk; c++) {
                    int lo = (int)((int64_t)angs.size() * c / nchunk), hi = (int)((int64_t)angs.size() * (c + 1) / nchunk);
                    Bundle b{};
                    b.np = hi - lo;
                    for (int i = lo; i < hi; i++)
                        b.ang[i - lo] = angs[i];
                    for (int i = b.np; i < kMaxPolar; i++)
                        b.ang[i] = angs[lo];
                    out.push_back(b);
                    out_phase.push_back(oct);
                    out_geom.push_back(geom);
                }
            }
        }
    };
    make_bundles(false, bundles, bundle_phase, bundle_geom);
    make_bundles(true, tbundles, tbundle_phase, tbundle_geom);
    // angle-family range of this handle: bundles outside it are dropped (a bundle never straddles families)
    {
        std::vector<int> family;
        h->n_family = angle_families(p, family);
        if (opt.family_begin != 0 || opt.family_end != 0) {
            if (opt.family_begin < 0 || opt.family_end > h->n_family || opt.family_begin >= opt.family_end)
                return fail(h, MOCB200_ERR_INVALID, "bad angle-family range [%d, %d) of %d", opt.family_begin,
                            opt.family_end, h->n_family);
            h->family_partial = opt.family_end - opt.family_begin < h->n_family;
            auto keep = [&](std::vector<Bundle> &bs, std::vector<int> &ph, std::vector<int> &ge) {
                size_t w = 0;
                for (size_t b = 0; b < bs.size(); b++) {
                    const int f = family[bs[b].ang[0]];
                    for (int q = 1; q < bs[b].np; q++)
                        if (family[bs[b].ang[q]] != f)
                            return false;
                    if (f >= opt.family_begin && f < opt.family_end)
                        bs[w] = bs[b], ph[w] = ph[b], ge[w] = ge[b], w++;
                }
                bs.resize(w), ph.resize(w), ge.resize(w);
                return true;
            };
            if (!keep(bundles, bundle_phase, bundle_geom) || !keep(tbundles, tbundle_phase, tbundle_geom))
                return fail(h, MOCB200_ERR_INVALID, "a polar bundle straddles two angle families");
        }
    }
    if (h->kernel == MOCB200_KERNEL_RCHUNK) { // one lane per polar angle: bundles of 1, 2 or 4; 8-bit flags + angle in 32 bits
        bool ok = p.n_ang < (1 << 23);
        for (const auto &b : tbundles)
            ok = ok && b.np != 3;
        if (!ok)
            h->kernel = MOCB200_KERNEL_CHUNK;
    }

    // ---- crossing lists, one pair per track ----
    std::vector<Cross> cross;
    std::vector<int32_t> cross_begin_fw(p.n_trk), cross_n_fw(p.n_trk), cross_begin_bw(p.n_trk), cross_n_bw(p.n_trk);
    {
        std::vector<Cross> fw, bw;
        for (int64_t t = 0; t < p.n_trk; t++) {
            fw.clear();
            bw.clear();
            int nseg = (int)(p.trk_seg_begin[t + 1] - p.trk_seg_begin[t]);
            build_crossings(p, t, nseg, fw, bw);
            cross_begin_fw[t] = (int32_t)cross.size();
            cross_n_fw[t]     = (int32_t)fw.size();
            cross.insert(cross.end(), fw.begin(), fw.end());
            cross_begin_bw[t] = (int32_t)cross.size();
            cross_n_bw[t]     = (int32_t)bw.size();
            cross.insert(cross.end(), bw.begin(), bw.end());
        }
    }

    // ---- work lists ----
    const bool jacobi = opt.boundary_update == MOCB200_BOUNDARY_JACOBI;
    for (int u = 0; u < p.n_unique; u++) {
        std::vector<int32_t> planes;
        for (int ip = h->plane_begin; ip < h->plane_end; ip++)
            if (p.plane_unique[ip] == u)
                planes.push_back(ip);
        if (planes.empty())
            continue;
        for (int phase = 0; phase < (jacobi ? 1 : 2); phase++) {
            for (int np = 1; np <= kMaxPolar; np++) {
                std::vector<Item> items;
                for (size_t b = 0; b < bundles.size(); b++) {
                    if (bundles[b].np != np || (!jacobi && bundle_phase[b] != phase))
                        continue;
                    int64_t t0 = p.geom_trk_begin[(size_t)u * p.n_geom + bundle_geom[b]];
                    int64_t t1 = p.geom_trk_begin[(size_t)u * p.n_geom + bundle_geom[b] + 1];
                    for (int64_t t = t0; t < t1; t++) {
                        int64_t s0 = p.trk_seg_begin[t];
                        int nseg   = (int)(p.trk_seg_begin[t + 1] - s0);
                        Item f{(int32_t)s0, nseg, p.trk_bc[2 * t], p.trk_bc[2 * t + 1], (int32_t)b, 0,
                               cross_begin_fw[t], cross_n_fw[t]};
                        Item r{(int32_t)(s0 + nseg - 1), nseg, p.trk_bc[2 * t + 1], p.trk_bc[2 * t], (int32_t)b, 1,
                               cross_begin_bw[t], cross_n_bw[t]};
                        items.push_back(f);
                        items.push_back(r);
                    }
                }
                if (items.empty())
                    continue;
                // longest first: dynamic scheduling then behaves like LPT and warps stay homogeneous
                std::stable_sort(items.begin(), items.end(), [](const Item &x, const Item &y) { return x.nseg > y.nseg; });
                WorkList wl;
                wl.unique = u, wl.phase = phase, wl.np = np;
                wl.n_items  = (int32_t)items.size();
                wl.n_planes = (int32_t)planes.size();
                int64_t segs = 0;
                for (const auto &it : items)
                    segs += it.nseg;
                wl.segs = segs * np * wl.n_planes;
                if ((int64_t)wl.n_items * p.n_group * wl.n_planes >= (int64_t)UINT32_MAX - 64)
                    return fail(h, MOCB200_ERR_INVALID, "work list too large for 32-bit scheduling");
                int rc = dev_upload(h, &wl.d_items, items);
                if (rc)
                    return rc;
                rc = dev_upload(h, &wl.d_planes, planes);
                if (rc)
                    return rc;
                h->lists.push_back(wl);
                h->stats.items[phase] += (int64_t)wl.n_items * wl.n_planes;
            }
        }
    }
    int64_t segs_ref = 0; // S: per-polar copies counted, both directions folded (= segs/2)
    for (const auto &wl : h->lists)
        segs_ref += wl.segs;
    h->stats.segments_per_sweep = segs_ref / 2;
    h->stats.unique_segments    = p.n_seg;

    // ---- track kernel: padded geometry, crossing lists with sentinels, per-4-segment crossing pointers ----
    {
        std::vector<int64_t> pbegin(p.n_trk + 1, 0);
        for (int64_t t = 0; t < p.n_trk; t++) {
            int64_t n = p.trk_seg_begin[t + 1] - p.trk_seg_begin[t];
            if (n < 1)
                return fail(h, MOCB200_ERR_INVALID, "track %lld has no segments", (long long)t);
            h->max_nseg   = std::max<int>(h->max_nseg, (int)n);
            pbegin[t + 1] = pbegin[t] + ((n + 3) & ~(int64_t)3);
        }
        const int64_t n_pseg = pbegin[p.n_trk];
        if (n_pseg >= (int64_t)INT32_MAX)
            return fail(h, MOCB200_ERR_INVALID, "padded segment count exceeds 32-bit indexing");
        std::vector<double> plen((size_t)n_pseg, 0.0);
        std::vector<int32_t> pfsr((size_t)n_pseg, 0);
        std::vector<int2> xptr((size_t)(n_pseg / 4));
        std::vector<Cross> xcross;
        std::vector<int32_t> xcross_f0((size_t)p.n_trk, 0); // per track: start of its forward list (backward follows)
        xcross.reserve(cross.size() + 2 * (size_t)p.n_trk);
        const Cross sentinel{INT32_MAX, 0};
        for (int64_t t = 0; t < p.n_trk; t++) {
            const int64_t s0 = p.trk_seg_begin[t];
            const int nseg   = (int)(p.trk_seg_begin[t + 1] - s0);
            std::copy(p.seg_len + s0, p.seg_len + s0 + nseg, plen.begin() + pbegin[t]);
            std::copy(p.seg_fsr + s0, p.seg_fsr + s0 + nseg, pfsr.begin() + pbegin[t]);
            const int32_t f0 = (int32_t)xcross.size();
            xcross_f0[t]     = f0;
            xcross.insert(xcross.end(), cross.begin() + cross_begin_fw[t], cross.begin() + cross_begin_fw[t] + cross_n_fw[t]);
            xcross.push_back(sentinel);
            const int32_t b0 = (int32_t)xcross.size();
            xcross.insert(xcross.end(), cross.begin() + cross_begin_bw[t], cross.begin() + cross_begin_bw[t] + cross_n_bw[t]);
            xcross.push_back(sentinel);
            // first crossing at or after the first node a 4-segment group owns, in either walk order
            int cf = f0, cb_hi = b0 + cross_n_bw[t]; // bwd pointers are found from the far end
            std::vector<int32_t> bw_first((size_t)(nseg + 3) / 4);
            for (int k0 = ((nseg - 1) / 4) * 4; k0 >= 0; k0 -= 4) {
                const int nb_min = std::max(nseg - k0 - 4, 0);
                // advance from the low end (nb ascending as k0 descends is NOT monotone here, so search)
                int lo = b0, hi = cb_hi;
                while (lo < hi) {
                    int mid = (lo + hi) / 2;
                    if (xcross[mid].node < nb_min)
                        lo = mid + 1;
                    else
                        hi = mid;
                }
                bw_first[k0 / 4] = lo;
            }
            for (int k0 = 0; k0 < nseg; k0 += 4) {
                while (xcross[cf].node < k0)
                    cf++;
                xptr[(size_t)(pbegin[t] + k0) / 4] = make_int2(cf, bw_first[k0 / 4]);
            }
        }
        for (int i = 0; i < 6; i++) // the kernels read ahead of the entry they test (RCHUNK: up to three past a sentinel)
            xcross.push_back(sentinel);
        int rc2;
        if ((rc2 = dev_upload(h, &h->d_pseg_len, plen)) || (rc2 = dev_upload(h, &h->d_pseg_fsr, pfsr)) ||
            (rc2 = dev_upload(h, &h->d_xptr, xptr)) || (rc2 = dev_upload(h, &h->d_xcross, xcross)))
            return rc2;

        // register-chunk kernel: regrouped FSR numbering (per unique plane) and the padded plane starts of its
        // q-bar / tally layout (every plane starts on a sector boundary)
        std::vector<std::vector<int32_t>> rc_lperm(p.n_unique);
        std::vector<int32_t> rc_plane_start(p.n_plane + 1, 0);
        if (h->kernel == MOCB200_KERNEL_RCHUNK) {
            std::vector<int32_t> perm((size_t)p.n_reg, 0);
            for (int ip = 0; ip < p.n_plane; ip++) {
                const int nreg_pl = (ip + 1 < p.n_plane ? p.plane_first_reg[ip + 1] : p.n_reg) - p.plane_first_reg[ip];
                rc_plane_start[ip + 1] = rc_plane_start[ip] + ((nreg_pl + 3) & ~3);
                const int u = p.plane_unique[ip];
                if (ip < h->plane_begin || ip >= h->plane_end) { // not swept by this handle: any one-to-one placement
                    for (int r = 0; r < nreg_pl; r++)
                        perm[(size_t)p.plane_first_reg[ip] + r] = rc_plane_start[ip] + r;
                    continue;
                }
                if (rc_lperm[u].empty()) {
                    static const char *no_group = getenv("MOCB200_RC_NOGROUP"); // tuning hook: keep the reference numbering
                    if (no_group && no_group[0] == '1') {
                        rc_lperm[u].resize(nreg_pl);
                        for (int r = 0; r < nreg_pl; r++)
                            rc_lperm[u][r] = r;
                    } else {
                        rc_lperm[u] = build_fsr_groups(p, u, nreg_pl);
                    }
                }
                if ((int)rc_lperm[u].size() != nreg_pl)
                    return fail(h, MOCB200_ERR_INVALID, "planes of one unique geometry differ in their FSR count");
                for (int r = 0; r < nreg_pl; r++)
                    perm[(size_t)p.plane_first_reg[ip] + r] = rc_plane_start[ip] + rc_lperm[u][r];
            }
            h->n_regp = rc_plane_start[p.n_plane];
            if ((rc2 = dev_upload(h, &h->d_fsr_perm, perm)))
                return rc2;
        }
        for (int u = 0; u < p.n_unique; u++) {
            std::vector<int32_t> planes;
            for (int ip = h->plane_begin; ip < h->plane_end; ip++)
                if (p.plane_unique[ip] == u)
                    planes.push_back(ip);
            if (planes.empty())
                continue;
            for (int phase = 0; phase < (jacobi ? 1 : 2); phase++) {
                for (int np = 1; np <= kMaxPolar; np++) {
                    struct UnitRec {
                        TrackUnit tu;
                        int4 len_begin; // per polar angle: where that angle's own segment lengths start
                        int64_t trk;    // representative track (crossing lists)
                    };
                    std::vector<UnitRec> recs;
                    for (size_t b = 0; b < tbundles.size(); b++) {
                        if (tbundles[b].np != np || (!jacobi && tbundle_phase[b] != phase))
                            continue;
                        int64_t t0 = p.geom_trk_begin[(size_t)u * p.n_geom + tbundle_geom[b]];
                        int64_t t1 = p.geom_trk_begin[(size_t)u * p.n_geom + tbundle_geom[b] + 1];
                        for (int64_t t = t0; t < t1; t++) {
                            UnitRec r{};
                            r.tu.seg_begin = (int32_t)pbegin[t];
                            r.tu.nseg      = (int32_t)(p.trk_seg_begin[t + 1] - p.trk_seg_begin[t]);
                            r.tu.bc0 = p.trk_bc[2 * t], r.tu.bc1 = p.trk_bc[2 * t + 1];
                            r.tu.bundle = (int32_t)b;
                            int32_t lb[kMaxPolar];
                            for (int q = 0; q < kMaxPolar; q++) {
                                const int gq = p.ang_geom[tbundles[b].ang[q]];
                                lb[q] = (int32_t)pbegin[p.geom_trk_begin[(size_t)u * p.n_geom + gq] + (t - t0)];
                            }
                            r.len_begin = make_int4(lb[0], lb[1], lb[2], lb[3]);
                            r.trk       = t;
                            recs.push_back(r);
                        }
                    }
                    if (recs.empty())
                        continue;
                    std::stable_sort(recs.begin(), recs.end(),
                                     [](const UnitRec &x, const UnitRec &y) { return x.tu.nseg > y.tu.nseg; });
                    std::vector<TrackUnit> units;
                    std::vector<int4> len_begin;
                    std::vector<int64_t> unit_trk;
                    for (const auto &r : recs) {
                        units.push_back(r.tu);
                        len_begin.push_back(r.len_begin);
                        unit_trk.push_back(r.trk);
                    }
                    TrackList tl;
                    for (auto &tu : units) {
                        tu.cpos = (int32_t)tl.pseg;
                        tl.pseg += (tu.nseg + 3) & ~3;
                    }
                    tl.unique = u, tl.phase = phase, tl.np = np;
                    tl.n_units  = (int32_t)units.size();
                    tl.max_nseg = units.front().nseg;
                    for (const auto &tu : units)
                        tl.nseg_desc.push_back(tu.nseg);
                    tl.n_planes = (int32_t)planes.size();
                    int64_t segs = 0;
                    for (const auto &tu : units)
                        segs += tu.nseg;
                    tl.segs = segs * np * tl.n_planes;
                    if ((int64_t)tl.n_units * tl.n_planes * p.n_group >= (int64_t)UINT32_MAX - 64)
                        return fail(h, MOCB200_ERR_INVALID, "track list too large for 32-bit scheduling");
                    if ((rc2 = dev_upload(h, &tl.d_units, units)) || (rc2 = dev_upload(h, &tl.d_planes, planes)) ||
                        (rc2 = dev_upload(h, &tl.d_len_begin, len_begin)))
                        return rc2;
                    {
                        // chunk-kernel descriptors: boundary linkage resolved once (boundary_condition.cpp:155-191)
                        std::vector<ChunkUnit> cu(units.size());
                        for (size_t i = 0; i < units.size(); i++) {
                            const TrackUnit &tu = units[i];
                            ChunkUnit &c = cu[i];
                            c.seg_begin = tu.seg_begin, c.nseg = tu.nseg, c.cpos = tu.cpos;
                            c.cross_begin = xcross_f0[unit_trk[i]];
                            c.n_fw = cross_n_fw[unit_trk[i]], c.n_bw = cross_n_bw[unit_trk[i]];
                            c.pad0 = c.pad1 = 0;
                            for (int q = 0; q < kMaxPolar; q++) {
                                const int ang = tbundles[tu.bundle].ang[q];
                                c.ang[q]  = ang;
                                c.in_f[q] = p.bc_offset[ang] + tu.bc0;
                                c.in_b[q] = p.bc_offset[ang + p.n_ang] + tu.bc1;
                                for (int dir = 0; dir < 2; dir++) {
                                    const int ao = ang + dir * p.n_ang, out_slot = dir ? tu.bc0 : tu.bc1;
                                    const int sx = p.bc_size_x[ao], face = out_slot >= sx ? 1 : 0;
                                    const int kind = p.bc_dst_kind[2 * ao + face];
                                    const int dst  = p.bc_dst_off[2 * ao + face] + out_slot - (face ? sx : 0);
                                    const int32_t enc = kind == 2 ? INT32_MIN : (kind == 1 ? dst : -(dst + 1));
                                    (dir ? c.out_b : c.out_f)[q] = enc;
                                }
                            }
                        }
                        std::vector<int2> pinfo;
                        for (int32_t ip : planes)
                            pinfo.push_back(make_int2(ip, p.plane_first_reg[ip]));
                        if ((rc2 = dev_upload(h, &tl.d_cunits, cu)) || (rc2 = dev_upload(h, &tl.d_pinfo, pinfo)))
                            return rc2;
                        if (h->kernel == MOCB200_KERNEL_RCHUNK) {
                            std::vector<int2> pinfo_rc;
                            for (int32_t ip : planes)
                                pinfo_rc.push_back(make_int2(ip, rc_plane_start[ip]));
                            if ((rc2 = build_rc_list(h, tl, cu, pfsr, rc_lperm[u], pinfo_rc, xcross)))
                                return rc2;
                        }
                    }
                    h->tlists.push_back(tl);
                }
            }
        }
        h->track_grid       = h->sm_count;
        h->scratch_per_warp = (h->max_nseg / 16 + 1) * 8 * kMaxPolar;
        for (const auto &tl : h->tlists)
            h->rc_scratch_per_team = std::max(h->rc_scratch_per_team, (tl.rc.max_nb + 1) * kMaxPolar);
        const size_t n_sc   = std::max((size_t)h->track_grid * (kWarpBlock / 32) * h->scratch_per_warp,
                                       (size_t)h->track_grid * 8 * h->rc_scratch_per_team);
        if ((rc2 = dev_alloc(h, &h->d_scratch, n_sc)))
            return rc2;
        if ((rc2 = dev_alloc(h, &h->d_gridbar, (size_t)4)))
            return rc2;
        {
            // mocb200_options.persistent, MOCB200_RC_PERSIST=0|1|2 overrides (A/B hook). Needs every CTA of the grid
            // resident at once (cooperative launch), which one CTA per SM always is.
            const char *pm  = getenv("MOCB200_RC_PERSIST");
            int coop        = 0;
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
            const int want  = pm ? atoi(pm) : opt.persistent;
            h->persist_mode = coop ? std::max(0, std::min(2, want)) : 0;
            const char *il   = getenv("MOCB200_RC_INTERLEAVE");
            h->rc_interleave = il ? atoi(il) : 1;
        }
    }

    // ---- 2D3D correction-factor tables ----
    h->n_cell_plane = p.n_cell_plane, h->n_geom = p.n_geom;
    h->have_sn_xs.assign(p.n_group, false);
    // A geometry whose ray data attribute one FSR to two coarse cells (rays through cell corners at very fine
    // spacings do) cannot use the per-FSR form of the correction sums: the handle then works for everything
    // except MOCB200_TALLY_CORRECTIONS, which reports why.
    constexpr int kCorrUnavailable = -1000;
    auto build_corr_tables = [&]() -> int {
        int rc3;
        // FSRs of every unique plane, their coarse cell (as the ray data attributes segments to cells,
        // correction_worker.hpp:145-162) and the path length of every geometry class through every FSR
        std::vector<int32_t> uniq_reg_begin(p.n_unique + 1, 0);
        std::vector<int> nreg_u(p.n_unique, 0);
        for (int u = 0; u < p.n_unique; u++) {
            int mx = -1;
            for (int gi = 0; gi < p.n_geom; gi++)
                for (int64_t t = p.geom_trk_begin[(size_t)u * p.n_geom + gi]; t < p.geom_trk_begin[(size_t)u * p.n_geom + gi + 1]; t++)
                    for (int64_t sidx = p.trk_seg_begin[t]; sidx < p.trk_seg_begin[t + 1]; sidx++)
                        mx = std::max(mx, (int)p.seg_fsr[sidx]);
            nreg_u[u]             = mx + 1;
            uniq_reg_begin[u + 1] = uniq_reg_begin[u] + nreg_u[u];
        }
        std::vector<int32_t> fsr_cell(uniq_reg_begin[p.n_unique], -1);
        std::vector<double> geom_len((size_t)uniq_reg_begin[p.n_unique] * p.n_geom, 0.0);
        for (int u = 0; u < p.n_unique; u++) {
            for (int gi = 0; gi < p.n_geom; gi++) {
                double *L = geom_len.data() + (size_t)uniq_reg_begin[u] * p.n_geom + (size_t)gi * nreg_u[u];
                for (int64_t t = p.geom_trk_begin[(size_t)u * p.n_geom + gi]; t < p.geom_trk_begin[(size_t)u * p.n_geom + gi + 1]; t++) {
                    const int64_t s0 = p.trk_seg_begin[t];
                    const int nseg   = (int)(p.trk_seg_begin[t + 1] - s0);
                    for (int k = 0; k < nseg; k++)
                        L[p.seg_fsr[s0 + k]] += p.seg_len[s0 + k];
                    int cell = p.trk_cm_start[4 * t + 0], iseg = 0;
                    for (int64_t k = p.trk_cm_begin[t]; k < p.trk_cm_begin[t + 1]; k++) {
                        const uint32_t c = p.cm_data[k];
                        const int s_fw = c & 0xF, n_fw = (c >> 8) & 0xFF;
                        if (s_fw != 7) {
                            for (int i = 0; i < n_fw; i++, iseg++) {
                                int32_t &fc = fsr_cell[uniq_reg_begin[u] + p.seg_fsr[s0 + iseg]];
                                if (fc >= 0 && fc != cell) {
                                    h->corr_why = "the ray data attribute an FSR to two coarse cells";
                                    return kCorrUnavailable;
                                }
                                fc = cell;
                            }
                        }
                        if (s_fw < 4) {
                            const int nb = p.coarse_nbr[4 * cell + s_fw];
                            cell         = nb < 0 ? cell : nb;
                        }
                    }
                }
            }
        }
        std::vector<int32_t> cell_fsr_begin((size_t)p.n_unique * (p.n_cell_plane + 1), 0), cell_fsr(fsr_cell.size(), 0);
        for (int u = 0; u < p.n_unique; u++) {
            int32_t *cb = cell_fsr_begin.data() + (size_t)u * (p.n_cell_plane + 1);
            for (int r = 0; r < nreg_u[u]; r++) {
                const int c = fsr_cell[uniq_reg_begin[u] + r];
                if (c < 0 || c >= p.n_cell_plane) {
                    h->corr_why = "an FSR is crossed by no ray";
                    return kCorrUnavailable;
                }
                cb[c + 1]++;
            }
            for (int c = 0; c < p.n_cell_plane; c++)
                cb[c + 1] += cb[c];
            std::vector<int32_t> fill(cb, cb + p.n_cell_plane);
            for (int r = 0; r < nreg_u[u]; r++)
                cell_fsr[uniq_reg_begin[u] + fill[fsr_cell[uniq_reg_begin[u] + r]]++] = r;
        }
        std::vector<int32_t> all_planes;
        for (int ip = h->plane_begin; ip < h->plane_end; ip++)
            all_planes.push_back(ip);
        if ((rc3 = dev_upload(h, &h->d_all_planes, all_planes)) ||
            (rc3 = dev_upload(h, &h->d_plane_unique, p.plane_unique, (size_t)p.n_plane)) ||
            (rc3 = dev_upload(h, &h->d_plane_cell_offset, p.plane_cell_offset, (size_t)p.n_plane)) ||
            (rc3 = dev_upload(h, &h->d_uniq_reg_begin, uniq_reg_begin)) ||
            (rc3 = dev_upload(h, &h->d_cell_fsr_begin, cell_fsr_begin)) || (rc3 = dev_upload(h, &h->d_cell_fsr, cell_fsr)) ||
            (rc3 = dev_upload(h, &h->d_geom_len, geom_len)) ||
            (rc3 = dev_upload(h, &h->d_ang_geom, p.ang_geom, (size_t)p.n_ang)) ||
            (rc3 = dev_upload(h, &h->d_coarse_surf, p.coarse_surf, (size_t)4 * p.n_cell_plane)) ||
            (rc3 = dev_upload(h, &h->d_area_x, p.ang_area_x, (size_t)p.n_ang)) ||
            (rc3 = dev_upload(h, &h->d_area_y, p.ang_area_y, (size_t)p.n_ang)) ||
            (rc3 = dev_upload(h, &h->d_ox, p.ang_ox, (size_t)p.n_ang)) ||
            (rc3 = dev_upload(h, &h->d_cell_dx, p.cell_dx, (size_t)p.n_cell_plane)) ||
            (rc3 = dev_upload(h, &h->d_cell_dy, p.cell_dy, (size_t)p.n_cell_plane)))
            return rc3;
        const size_t nsx = (size_t)p.n_plane * p.n_cell_plane * h->GP;
        if ((rc3 = dev_alloc(h, &h->d_sn_xs, nsx)))
            return rc3;
        CUDA_TRY(h, cudaMemset(h->d_sn_xs, 0, nsx * sizeof(double)));
        h->have_corr = true;
        return MOCB200_OK;
    };
    if (p.ang_area_x && p.ang_area_y && p.ang_ox && p.cell_dx && p.cell_dy) {
        const int rc3 = build_corr_tables();
        if (rc3 != MOCB200_OK && rc3 != kCorrUnavailable)
            return rc3;
    } else {
        h->corr_why = "the problem was created without the 2D3D correction tables";
    }

    // ---- uploads ----
    int rc;
#define UP(dst, src, n)                                                                                      \
    if ((rc = dev_upload(h, &(dst), (src), (size_t)(n))))                                                    \
        return rc;
    UP(h->d_seg_len, p.seg_len, p.n_seg);
    UP(h->d_seg_fsr, p.seg_fsr, p.n_seg);
    if ((rc = dev_upload(h, &h->d_cross, cross)))
        return rc;
    if ((rc = dev_upload(h, &h->d_bundles, bundles)) || (rc = dev_upload(h, &h->d_tbundles, tbundles)))
        return rc;
    UP(h->d_rsin, p.ang_rsintheta, p.n_ang);
    UP(h->d_wt, p.wt_v_st, (size_t)p.n_plane * p.n_ang);
    {
        std::vector<double> cw((size_t)p.n_plane * p.n_ang * 2), fw(cw.size());
        for (size_t i = 0; i < (size_t)p.n_plane * p.n_ang; i++) {
            cw[2 * i] = p.cur_wx[i], cw[2 * i + 1] = p.cur_wy[i];
            fw[2 * i] = p.flx_wx[i], fw[2 * i + 1] = p.flx_wy[i];
        }
        if ((rc = dev_upload(h, &h->d_curw, cw)))
            return rc;
        if ((rc = dev_upload(h, &h->d_flxw, fw)))
            return rc;
    }
    UP(h->d_bc_offset, p.bc_offset, 2 * p.n_ang);
    UP(h->d_bc_size_x, p.bc_size_x, 2 * p.n_ang);
    UP(h->d_bc_dst_off, p.bc_dst_off, 4 * p.n_ang);
    UP(h->d_bc_dst_kind, p.bc_dst_kind, 4 * p.n_ang);
    UP(h->d_plane_first_reg, p.plane_first_reg, p.n_plane);
    UP(h->d_plane_surf_offset, p.plane_surf_offset, p.n_plane);
    UP(h->d_vol, p.vol, p.n_reg);
    if ((rc = dev_alloc(h, &h->d_exp, (size_t)p.exp_n + 4))) // padded: the table is staged in 16-byte units
        return rc;
    CUDA_TRY(h, cudaMemset(h->d_exp, 0, ((size_t)p.exp_n + 4) * sizeof(double)));
    CUDA_TRY(h, cudaMemcpy(h->d_exp, p.exp_table, ((size_t)p.exp_n + 2) * sizeof(double), cudaMemcpyHostToDevice));
#undef UP

    const size_t nrg = (size_t)p.n_reg * h->GP;
    if ((rc = dev_alloc(h, &h->d_xq, nrg)))
        return rc;
    CUDA_TRY(h, cudaMemset(h->d_xq, 0, nrg * sizeof(double2)));
    double **fsr_arrays[] = {&h->d_xstr, &h->d_xstr_src, &h->d_xs_self, &h->d_src, &h->d_flux, &h->d_qbar, &h->d_tally};
    for (double **arr : fsr_arrays) {
        if ((rc = dev_alloc(h, arr, nrg)))
            return rc;
        CUDA_TRY(h, cudaMemset(*arr, 0, nrg * sizeof(double)));
    }
    const size_t nbc = (size_t)p.n_plane * p.bc_per_group * h->GP;
    for (int i = 0; i < (jacobi ? 2 : 1); i++) {
        if ((rc = dev_alloc(h, &h->d_bc[i], nbc)))
            return rc;
        CUDA_TRY(h, cudaMemset(h->d_bc[i], 0, nbc * sizeof(double)));
    }
    if (!jacobi)
        h->d_bc[1] = h->d_bc[0];
    const size_t nsf = (size_t)p.n_surf * h->GP;
    if ((rc = dev_alloc(h, &h->d_current, nsf)) || (rc = dev_alloc(h, &h->d_surfflux, nsf)))
        return rc;
    CUDA_TRY(h, cudaMemset(h->d_current, 0, nsf * sizeof(double)));
    CUDA_TRY(h, cudaMemset(h->d_surfflux, 0, nsf * sizeof(double)));

    h->stage_elems = std::max<size_t>({(size_t)p.n_reg, (size_t)p.bc_per_group, (size_t)p.n_surf,
                                       (size_t)p.n_plane * p.n_cell_plane}) * p.n_group;
    if ((rc = dev_alloc(h, &h->d_stage, h->stage_elems)))
        return rc;
    CUDA_TRY(h, cudaMallocHost((void **)&h->h_stage, h->stage_elems * sizeof(double)));
    h->io_elems = 2 * (size_t)p.n_reg + (size_t)(h->plane_end - h->plane_begin) * p.bc_per_group + 2 * (size_t)p.n_surf;
    if ((rc = dev_alloc(h, &h->d_in, h->io_elems)) || (rc = dev_alloc(h, &h->d_out, h->io_elems)))
        return rc;
    CUDA_TRY(h, cudaMallocHost((void **)&h->h_in, h->io_elems * sizeof(double)));
    CUDA_TRY(h, cudaMallocHost((void **)&h->h_out, h->io_elems * sizeof(double)));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
    h->n_counters = (int)(h->lists.size() + h->tlists.size());
    if ((rc = dev_alloc(h, &h->d_counters, (size_t)h->n_counters)))
        return rc;
    h->have_xs.assign(p.n_group, false);
    h->cache_valid.assign(p.n_group, false);
    const size_t n_qg = (size_t)p.n_group * std::max(p.n_reg, h->n_regp);
    if ((rc = dev_alloc(h, &h->d_qg, n_qg)) || (rc = dev_alloc(h, &h->d_tg, n_qg)))
        return rc;
    h->stats.device_bytes = h->device_bytes;
    h->stats.kernel       = h->kernel;
    for (const auto &tl : h->tlists)
        h->stats.swept_segments += tl.segs / std::max(1, tl.np) / std::max(1, tl.n_planes);
    return MOCB200_OK;
}

typedef void (*SweepFn)(const SweepArgs);

SweepFn pick_kernel(int np, int tally)
{
    switch (np * 2 + (tally ? 1 : 0)) {
    case 2: return sweep_kernel<1, 0>;
    case 3: return sweep_kernel<1, 1>;
    case 4: return sweep_kernel<2, 0>;
    case 5: return sweep_kernel<2, 1>;
    case 6: return sweep_kernel<3, 0>;
    case 7: return sweep_kernel<3, 1>;
    case 8: return sweep_kernel<4, 0>;
    case 9: return sweep_kernel<4, 1>;
    }
    return nullptr;
}

typedef void (*CacheFn)(const CacheArgs);

CacheFn pick_cache_kernel(int np)
{
    switch (np) {
    case 1: return exp_cache_kernel<1>;
    case 2: return exp_cache_kernel<2>;
    case 3: return exp_cache_kernel<3>;
    case 4: return exp_cache_kernel<4>;
    }
    return nullptr;
}

typedef void (*WarpFn)(const WarpArgs);

template <int GL, bool CACHED> WarpFn pick_warp_gl(int np, int tally)
{
    switch (np * 3 + tally) {
    case 3: return sweep_warp_kernel<GL, 1, 0, CACHED>;
    case 4: return sweep_warp_kernel<GL, 1, 1, CACHED>;
    case 5: return sweep_warp_kernel<GL, 1, 2, CACHED>;
    case 6: return sweep_warp_kernel<GL, 2, 0, CACHED>;
    case 7: return sweep_warp_kernel<GL, 2, 1, CACHED>;
    case 8: return sweep_warp_kernel<GL, 2, 2, CACHED>;
    case 9: return sweep_warp_kernel<GL, 3, 0, CACHED>;
    case 10: return sweep_warp_kernel<GL, 3, 1, CACHED>;
    case 11: return sweep_warp_kernel<GL, 3, 2, CACHED>;
    case 12: return sweep_warp_kernel<GL, 4, 0, CACHED>;
    case 13: return sweep_warp_kernel<GL, 4, 1, CACHED>;
    case 14: return sweep_warp_kernel<GL, 4, 2, CACHED>;
    }
    return nullptr;
}

WarpFn pick_warp_kernel(int gl, int np, int tally, bool cached)
{
    if (cached)
        return gl == 1 ? pick_warp_gl<1, true>(np, tally) : pick_warp_gl<8, true>(np, tally);
    return gl == 1 ? pick_warp_gl<1, false>(np, tally) : pick_warp_gl<8, false>(np, tally);
}

typedef void (*ChunkFn)(const ChunkArgs);

template <int NW, int TALLY> ChunkFn pick_chunk_np(int np)
{
    switch (np) {
    case 1: return sweep_chunk_kernel<1, NW, TALLY>;
    case 2: return sweep_chunk_kernel<2, NW, TALLY>;
    case 3: return sweep_chunk_kernel<3, NW, TALLY>;
    case 4: return sweep_chunk_kernel<4, NW, TALLY>;
    }
    return nullptr;
}

ChunkFn pick_chunk_kernel(int np, int nw, int tally)
{
    switch ((nw - 1) * 3 + tally) {
    case 0: return pick_chunk_np<1, 0>(np);
    case 1: return pick_chunk_np<1, 1>(np);
    case 2: return pick_chunk_np<1, 2>(np);
    case 3: return pick_chunk_np<2, 0>(np);
    case 4: return pick_chunk_np<2, 1>(np);
    case 5: return pick_chunk_np<2, 2>(np);
    }
    return nullptr;
}

typedef void (*RcCacheFn)(const RcCacheArgs);

// the instantiations live in moc_rc_p{1,2,4}.cu (one translation unit per lane count, compiled in parallel)
// the tallying sweeps need more registers per thread: one team less per CTA when that instantiation exists
RcConfig rc_tally_config(int np, int tally, const RcConfig &c);

RcPersistFn pick_rc_persist_kernel(int np, const RcConfig &c)
{
    switch (np) {
    case 1: return pick_rc_persist_kernel_p1(c);
    case 2: return pick_rc_persist_kernel_p2(c);
    case 4: return pick_rc_persist_kernel_p4(c);
    }
    return nullptr;
}

RcFn pick_rc_kernel(int np, int tally, const RcConfig &c)
{
    switch (np) {
    case 1: return pick_rc_kernel_p1(tally, c);
    case 2: return pick_rc_kernel_p2(tally, c);
    case 4: return pick_rc_kernel_p4(tally, c);
    }
    return nullptr;
}

RcConfig rc_tally_config(int np, int tally, const RcConfig &c)
{
    static const char *tt = getenv("MOCB200_RC_TALLY_TEAMS"); // A/B hook: teams per CTA of the tallying variants
    if (tally != 0 && tt && atoi(tt) >= 1 && pick_rc_kernel(np, tally, RcConfig{c.LMAX, c.NW, atoi(tt)}))
        return RcConfig{c.LMAX, c.NW, atoi(tt)};
    if (tally == 0 || c.TEAMS <= 3) // three four-warp teams leave 168 registers per thread
        return c;
    const RcConfig t{c.LMAX, c.NW, c.TEAMS - 1};
    return (t.TEAMS >= 1 && pick_rc_kernel(np, tally, t)) ? t : c;
}

RcCacheFn pick_rc_cache_kernel(int np)
{
    switch (np) {
    case 1: return rc_cache_kernel<1>;
    case 2: return rc_cache_kernel<2>;
    case 4: return rc_cache_kernel<4>;
    }
    return nullptr;
}

constexpr int kChunkSmemBudget = 232448 - 12800; // opt-in dynamic shared memory per CTA minus the static part

// launch geometry of the chunk kernel for a list: segments a team stages at once, warps per team,
// teams per CTA
void chunk_geometry(const std::vector<int32_t> &nseg_desc, int np, int cap_opt, bool tally, int *caps, int *nw,
                    int *teams)
{
    const int max_nseg  = nseg_desc.empty() ? 32 : nseg_desc.front();
    const int per_seg   = 8 * (np + 2) + 8 + (tally ? 4 : 0);
    const int cap_limit = ((kChunkSmemBudget / 4) / per_seg) & ~31; // at least 4 tracks in flight per SM
    int c = (max_nseg + 31) & ~31;
    c = std::min(c, cap_limit);
    // Shared memory bounds the tracks in flight per SM. Staging the longest tracks whole is not always best: a
    // smaller cap lets more teams run and only the few tracks above it pay a second pass (super-blocks).
    // Model: time ~ (1 + share of the segments in tracks above the cap) / teams^0.75 (measured on C5G7-2D:
    // cap 768 / 7 teams 0.1392 ms, 672 / 8 teams 0.1362, 640 / 8 0.143, 576 / 9 0.152).
    if (c >= 256 && cap_opt == 0) {
        const int max_teams = std::min(chunk_max_warps(2) / 2, kChunkMaxTeams);
        double total = 0.0;
        for (int32_t n : nseg_desc)
            total += n;
        auto teams_of = [&](int cap) { return std::min<int>(max_teams, kChunkSmemBudget / (int)chunk_warp_bytes(cap, np, tally)); };
        auto model = [&](int cap, int k) {
            double above = 0.0;
            for (int32_t n : nseg_desc) {
                if (n <= cap)
                    break;
                above += n;
            }
            return (1.0 + above / std::max(total, 1.0)) / std::pow((double)std::max(1, k), 0.75);
        };
        double best = model(c, teams_of(c));
        int best_c  = c;
        for (int k = teams_of(c) + 1; k <= max_teams; k++) {
            const int cap = ((kChunkSmemBudget / k) / per_seg) & ~31; // largest cap that lets k teams run
            if (cap < 256)
                break;
            const double t = model(cap, k);
            if (t < best)
                best = t, best_c = cap;
        }
        c = best_c;
    }
    if (cap_opt > 0)
        c = std::min(c, (cap_opt + 31) & ~31);
    static const char *force_cap = getenv("MOCB200_CHUNK_CAP"); // tuning hook: staging cap in segments
    if (force_cap && atoi(force_cap) > 0)
        c = std::min(c, (atoi(force_cap) + 31) & ~31);
    *caps  = c;
    *nw    = c >= 256 ? 2 : 1;
    static const char *force_nw = getenv("MOCB200_CHUNK_NW"); // tuning hook: warps per track (1 or 2)
    if (force_nw && force_nw[0] >= '1' && force_nw[0] <= '0' + kChunkMaxTeam)
        *nw = force_nw[0] - '0';
    *teams = std::max(1, std::min<int>(std::min(chunk_max_warps(*nw) / *nw, kChunkMaxTeams),
                                       kChunkSmemBudget / (int)chunk_warp_bytes(c, np, tally)));
}

int track_smem_bytes(const mocb200_sweeper *h)
{
    return (((h->exp_n + 2) * (int)sizeof(double)) + 15) & ~15;
}

// page-locked host memory (cudaHostAlloc / cudaHostRegister) can be the source or target of a DMA directly
bool is_pinned(const void *p)
{
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

int check_groups(mocb200_sweeper *h, int g_begin, int g_count)
{
    if (!h)
        return MOCB200_ERR_INVALID;
    if (g_begin < 0 || g_count < 1 || g_begin + g_count > h->G)
        return fail(h, MOCB200_ERR_INVALID, "group range [%d, %d) outside [0, %d)", g_begin, g_begin + g_count, h->G);
    return MOCB200_OK;
}

// host columns [g_count][n_total], rows [r0, r1) of every column -> device [n_total][GP]: what a handle that owns
// a plane range needs of a per-FSR array
int upload_column_rows(mocb200_sweeper *h, const double *host, int64_t n_total, int64_t r0, int64_t r1, int g_begin,
                       int g_count, double *dst)
{
    const int64_t nr = r1 - r0;
    for (int g = 0; g < g_count; g++)
        std::memcpy(h->h_stage + (size_t)g * nr, host + (size_t)g * n_total + r0, (size_t)nr * sizeof(double));
    const size_t elems = (size_t)nr * g_count;
    CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, h->h_stage, elems * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    scatter_columns_kernel<<<grid_for((int64_t)elems, 256, h->sm_count), 256, 0, h->stream>>>(nr, h->GP, g_begin, g_count,
                                                                                             h->d_stage, dst + (size_t)r0 * h->GP);
    h->stats.kernel_launches++;
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream)); // the pinned staging buffer is reused by the next call
    return MOCB200_OK;
}

// host columns [g_count][n] -> device [n][GP]
int upload_columns(mocb200_sweeper *h, const double *host, int64_t n, int g_begin, int g_count, double *dst)
{
    const size_t elems = (size_t)n * g_count;
    std::memcpy(h->h_stage, host, elems * sizeof(double));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_stage, h->h_stage, elems * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    scatter_columns_kernel<<<grid_for((int64_t)elems, 256, h->sm_count), 256, 0, h->stream>>>(n, h->GP, g_begin, g_count,
                                                                                             h->d_stage, dst);
    h->stats.kernel_launches++;
    CUDA_TRY(h, cudaGetLastError());
    // the pinned staging buffer is reused by the next call
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MOCB200_OK;
}

int download_columns(mocb200_sweeper *h, double *host, int64_t n, int g_begin, int g_count, const double *src)
{
    const size_t elems = (size_t)n * g_count;
    gather_columns_kernel<<<grid_for((int64_t)elems, 256, h->sm_count), 256, 0, h->stream>>>(n, h->GP, g_begin, g_count,
                                                                                            src, h->d_stage);
    h->stats.kernel_launches++;
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaMemcpyAsync(h->h_stage, h->d_stage, elems * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    std::memcpy(host, h->h_stage, elems * sizeof(double));
    return MOCB200_OK;
}

} // namespace

extern "C" {

const char *mocb200_version(void)
{
    return "mocc_b200 0.1 (sm_100a)";
}

const char *mocb200_last_error(const mocb200_sweeper *h)
{
    return h ? h->error.c_str() : g_create_error.c_str();
}

int mocb200_create(const mocb200_problem *prob, const mocb200_options *opt, mocb200_sweeper **out)
{
    if (!out)
        return fail(nullptr, MOCB200_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int rc = validate(prob);
    if (rc)
        return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, MOCB200_ERR_NO_DEVICE, "no CUDA device available (the MoC sweep has no CPU fallback)");
    mocb200_sweeper *h = new mocb200_sweeper();
    if (opt)
        h->opt = *opt;
    h->device = h->opt.device;
    if (h->device < 0 || h->device >= ndev) {
        const int bad = h->device;
        delete h;
        return fail(nullptr, MOCB200_ERR_INVALID, "device %d out of range (%d devices)", bad, ndev);
    }
    cudaError_t e = cudaSetDevice(h->device);
    if (e == cudaSuccess)
        e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess)
        e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess)
        e = cudaEventCreate(&h->ev1);
    if (e == cudaSuccess)
        e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
    if (e != cudaSuccess) {
        fail(nullptr, MOCB200_ERR_CUDA, "device set-up failed: %s", cudaGetErrorString(e));
        delete h;
        return MOCB200_ERR_CUDA;
    }
    h->stream = h->own_stream;
    rc = build(h, *prob);
    if (rc) {
        g_create_error = h->error;
        mocb200_destroy(h);
        return rc;
    }
    // opt in to the shared memory the table needs, for every kernel variant
    const int smem = (h->exp_n + 2) * (int)sizeof(double);
    for (int np = 1; np <= kMaxPolar; np++)
        for (int t = 0; t < 2; t++) {
            e = cudaFuncSetAttribute((const void *)pick_kernel(np, t), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) {
                fail(nullptr, MOCB200_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
                mocb200_destroy(h);
                return MOCB200_ERR_CUDA;
            }
        }
    for (int np = 1; np <= kMaxPolar && e == cudaSuccess; np++)
        for (int nw = 1; nw <= kChunkMaxTeam && e == cudaSuccess; nw++)
            for (int t = 0; t < 3 && e == cudaSuccess; t++)
                e = cudaFuncSetAttribute((const void *)pick_chunk_kernel(np, nw, t),
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kChunkSmemBudget);
    if (e == cudaSuccess)
        for (int np = 1; np <= kMaxPolar && e == cudaSuccess; np++)
            e = cudaFuncSetAttribute((const void *)pick_cache_kernel(np), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (h->kernel == MOCB200_KERNEL_RCHUNK) {
        for (const auto &tl : h->tlists) {
            const RcList &rc = tl.rc;
            const RcConfig cfg{rc.LMAX, rc.NW, rc.TEAMS};
            for (int t = 0; t < 3 && e == cudaSuccess; t++) {
                const RcConfig ct = rc_tally_config(rc.P, t, cfg);
                e = cudaFuncSetAttribute((const void *)pick_rc_kernel(rc.P, t, ct), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(ct.TEAMS * rc_team_bytes(rc.P, rc.LMAX, rc.NW)));
            }
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute((const void *)pick_rc_cache_kernel(rc.P), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        }
    }
    if (e != cudaSuccess) {
        fail(nullptr, MOCB200_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        mocb200_destroy(h);
        return MOCB200_ERR_CUDA;
    }
    for (int gl = 1; gl <= 8; gl *= 8)
        for (int np = 1; np <= kMaxPolar; np++)
            for (int t = 0; t < 3; t++) {
                e = cudaFuncSetAttribute((const void *)pick_warp_kernel(gl, np, t, false),
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, track_smem_bytes(h));
                if (e != cudaSuccess) {
                    fail(nullptr, MOCB200_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
                    mocb200_destroy(h);
                    return MOCB200_ERR_CUDA;
                }
            }
    *out = h;
    return MOCB200_OK;
}

int mocb200_destroy(mocb200_sweeper *h)
{
    if (!h)
        return MOCB200_OK;
    cudaSetDevice(h->device);
    if (h->own_stream)
        cudaStreamSynchronize(h->own_stream);
    for (void *p : h->allocs)
        cudaFree(p);
    for (auto &tl : h->tlists)
        if (tl.d_cache)
            cudaFree(tl.d_cache);
    for (double *q : {h->d_dsum, h->d_ssum, h->d_alpha, h->d_beta})
        if (q)
            cudaFree(q);
    if (h->h_stage)
        cudaFreeHost(h->h_stage);
    if (h->h_in)
        cudaFreeHost(h->h_in);
    if (h->h_out)
        cudaFreeHost(h->h_out);
    if (h->ev_in)
        cudaEventDestroy(h->ev_in);
    for (auto &pr : h->ev_pool) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    if (h->ev0)
        cudaEventDestroy(h->ev0);
    if (h->ev1)
        cudaEventDestroy(h->ev1);
    if (h->own_stream)
        cudaStreamDestroy(h->own_stream);
    delete h;
    return MOCB200_OK;
}

int mocb200_set_stream(mocb200_sweeper *h, void *cuda_stream)
{
    if (!h)
        return MOCB200_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return MOCB200_OK;
}

int mocb200_synchronize(mocb200_sweeper *h)
{
    if (!h)
        return MOCB200_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MOCB200_OK;
}

int mocb200_set_xs(mocb200_sweeper *h, int g_begin, int g_count, const double *xstr, const double *xstr_src,
                   const double *xs_self)
{
    int rc = check_groups(h, g_begin, g_count);
    if (rc)
        return rc;
    if (!xstr || !xs_self)
        return fail(h, MOCB200_ERR_INVALID, "xstr and xs_self are required");
    CUDA_TRY(h, cudaSetDevice(h->device));
    // only the FSR range of this handle's macroplanes (the host arrays keep their global indexing)
    if ((rc = upload_column_rows(h, xstr, h->n_reg, h->reg_lo, h->reg_hi, g_begin, g_count, h->d_xstr)))
        return rc;
    if ((rc = upload_column_rows(h, xstr_src ? xstr_src : xstr, h->n_reg, h->reg_lo, h->reg_hi, g_begin, g_count, h->d_xstr_src)))
        return rc;
    if ((rc = upload_column_rows(h, xs_self, h->n_reg, h->reg_lo, h->reg_hi, g_begin, g_count, h->d_xs_self)))
        return rc;
    for (int g = g_begin; g < g_begin + g_count; g++) {
        h->have_xs[g]     = true;
        h->cache_valid[g] = false;
    }
    return MOCB200_OK;
}

#define COLUMN_SETTER(NAME, FIELD)                                                                           \
    int NAME(mocb200_sweeper *h, int g_begin, int g_count, const double *v)                                  \
    {                                                                                                        \
        int rc = check_groups(h, g_begin, g_count);                                                          \
        if (rc)                                                                                              \
            return rc;                                                                                       \
        if (!v)                                                                                              \
            return fail(h, MOCB200_ERR_INVALID, #NAME ": NULL array");                                       \
        CUDA_TRY(h, cudaSetDevice(h->device));                                                               \
        return upload_columns(h, v, h->n_reg, g_begin, g_count, h->FIELD);                                   \
    }
COLUMN_SETTER(mocb200_set_source, d_src)
COLUMN_SETTER(mocb200_set_flux, d_flux)
COLUMN_SETTER(mocb200_set_qbar, d_qbar)
#undef COLUMN_SETTER

int mocb200_get_source(mocb200_sweeper *h, int g_begin, int g_count, double *src)
{
    int rc = check_groups(h, g_begin, g_count);
    if (rc)
        return rc;
    if (!src)
        return fail(h, MOCB200_ERR_INVALID, "get_source: NULL array");
    CUDA_TRY(h, cudaSetDevice(h->device));
    return download_columns(h, src, h->n_reg, g_begin, g_count, h->d_src);
}

int mocb200_set_source_xs(mocb200_sweeper *h, int n_mat, const int32_t *fsr_mat, const double *xsnf, const double *xsch,
                          const double *scat, const int32_t *scat_band)
{
    if (!h)
        return MOCB200_ERR_INVALID;
    if (n_mat < 1 || !fsr_mat || !xsnf || !xsch || !scat)
        return fail(h, MOCB200_ERR_INVALID, "set_source_xs: NULL table or no material");
    const int G = h->G;
    for (int r = 0; r < h->n_reg; r++)
        if (fsr_mat[r] < 0 || fsr_mat[r] >= n_mat)
            return fail(h, MOCB200_ERR_INVALID, "set_source_xs: FSR %d has cross-section region %d of %d", r, fsr_mat[r], n_mat);
    std::vector<int32_t> band((size_t)n_mat * G * 2);
    for (int m = 0; m < n_mat; m++)
        for (int g = 0; g < G; g++) {
            int lo = G, hi = -1; // ScatteringRow: first / last non-zero entry of the row (scattering_matrix.cpp:40-62)
            if (scat_band) {
                lo = scat_band[((size_t)m * G + g) * 2], hi = scat_band[((size_t)m * G + g) * 2 + 1];
                if (lo < 0 || hi >= G)
                    return fail(h, MOCB200_ERR_INVALID, "set_source_xs: scattering band [%d, %d] outside the groups", lo, hi);
            } else {
                for (int gg = 0; gg < G; gg++)
                    if (scat[((size_t)m * G + g) * G + gg] != 0.0)
                        lo = std::min(lo, gg), hi = std::max(hi, gg);
            }
            band[((size_t)m * G + g) * 2] = lo, band[((size_t)m * G + g) * 2 + 1] = hi;
        }
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    int rc;
    if (!h->d_fsr_mat && ((rc = dev_alloc(h, &h->d_fsr_mat, (size_t)h->n_reg)) || (rc = dev_alloc(h, &h->d_fs, (size_t)h->n_reg))))
        return rc;
    if (n_mat > h->n_mat) { // tables grow: the old ones stay in the handle's allocation list
        if ((rc = dev_alloc(h, &h->d_mat_nf, (size_t)n_mat * G)) || (rc = dev_alloc(h, &h->d_mat_ch, (size_t)n_mat * G)) ||
            (rc = dev_alloc(h, &h->d_mat_scat, (size_t)n_mat * G * G)) || (rc = dev_alloc(h, &h->d_scat_band, (size_t)n_mat * G * 2)))
            return rc;
    }
    h->n_mat = std::max(h->n_mat, n_mat);
    CUDA_TRY(h, cudaMemcpy(h->d_fsr_mat, fsr_mat, sizeof(int32_t) * h->n_reg, cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->d_mat_nf, xsnf, sizeof(double) * n_mat * G, cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->d_mat_ch, xsch, sizeof(double) * n_mat * G, cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->d_mat_scat, scat, sizeof(double) * n_mat * G * G, cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->d_scat_band, band.data(), sizeof(int32_t) * band.size(), cudaMemcpyHostToDevice));
    h->stats.device_bytes = h->device_bytes;
    return MOCB200_OK;
}

int mocb200_set_external_source(mocb200_sweeper *h, const double *ext)
{
    if (!h)
        return MOCB200_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (!ext) {
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        h->d_ext = nullptr; // stays in the allocation list
        return MOCB200_OK;
    }
    int rc;
    double *d = nullptr;
    if ((rc = dev_alloc(h, &d, (size_t)h->n_reg * h->GP)))
        return rc;
    CUDA_TRY(h, cudaMemsetAsync(d, 0, sizeof(double) * h->n_reg * h->GP, h->stream));
    if ((rc = upload_columns(h, ext, h->n_reg, 0, h->G, d)))
        return rc;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->d_ext = d;
    return MOCB200_OK;
}

int mocb200_fission_source(mocb200_sweeper *h, double k)
{
    if (!h)
        return MOCB200_ERR_INVALID;
    if (!h->d_fsr_mat)
        return fail(h, MOCB200_ERR_STATE, "mocb200_set_source_xs has not been called");
    if (!(k > 0.0))
        return fail(h, MOCB200_ERR_INVALID, "fission_source: k = %g", k);
    CUDA_TRY(h, cudaSetDevice(h->device));
    fission_source_kernel<<<grid_for(h->reg_hi - h->reg_lo, 256, h->sm_count), 256, 0, h->stream>>>(
        h->reg_lo, h->reg_hi, h->G, h->GP, 1.0 / k, h->d_fsr_mat, h->d_mat_nf, h->d_flux, h->d_fs);
    h->stats.kernel_launches++;
    h->have_fs = true;
    CUDA_TRY(h, cudaGetLastError());
    return MOCB200_OK;
}

int mocb200_set_fission_source(mocb200_sweeper *h, const double *fs)
{
    if (!h || !fs)
        return MOCB200_ERR_INVALID;
    if (!h->d_fsr_mat)
        return fail(h, MOCB200_ERR_STATE, "mocb200_set_source_xs has not been called");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaMemcpy(h->d_fs, fs, sizeof(double) * h->n_reg, cudaMemcpyHostToDevice));
    h->have_fs = true;
    return MOCB200_OK;
}

int mocb200_get_fission_source(mocb200_sweeper *h, double *fs)
{
    if (!h || !fs)
        return MOCB200_ERR_INVALID;
    if (!h->have_fs)
        return fail(h, MOCB200_ERR_STATE, "no fission source on the device yet");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaMemcpy(fs, h->d_fs, sizeof(double) * h->n_reg, cudaMemcpyDeviceToHost));
    return MOCB200_OK;
}

int mocb200_build_source(mocb200_sweeper *h, int g_begin, int g_count)
{
    int rc = check_groups(h, g_begin, g_count);
    if (rc)
        return rc;
    if (!h->d_fsr_mat)
        return fail(h, MOCB200_ERR_STATE, "mocb200_set_source_xs has not been called");
    if (!h->have_fs)
        return fail(h, MOCB200_ERR_STATE, "mocb200_fission_source / mocb200_set_fission_source has not been called");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int64_t n = (int64_t)(h->reg_hi - h->reg_lo) * g_count;
    group_source_kernel<<<grid_for(n, 256, h->sm_count), 256, 0, h->stream>>>(
        h->reg_lo, h->reg_hi, h->G, h->GP, g_begin, g_count, h->d_fsr_mat, h->d_mat_ch, h->d_mat_scat, h->d_scat_band,
        h->d_ext, h->d_fs, h->d_flux, h->d_src);
    h->stats.kernel_launches++;
    CUDA_TRY(h, cudaGetLastError());
    return MOCB200_OK;
}

int mocb200_get_flux(mocb200_sweeper *h, int g_begin, int g_count, double *flux)
{
    int rc = check_groups(h, g_begin, g_count);
    if (rc)
        return rc;
    if (!flux)
        return fail(h, MOCB200_ERR_INVALID, "get_flux: NULL array");
    CUDA_TRY(h, cudaSetDevice(h->device));
    return download_columns(h, flux, h->n_reg, g_begin, g_count, h->d_flux);
}

int mocb200_set_boundary(mocb200_sweeper *h, int plane, int g_begin, int g_count, const double *bc)
{
    int rc = check_groups(h, g_begin, g_count);
    if (rc)
        return rc;
    if (plane < 0 || plane >= h->n_plane || !bc)
        return fail(h, MOCB200_ERR_INVALID, "set_boundary: bad plane %d or NULL array", plane);
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t off = (size_t)plane * h->bcpg * h->GP;
    if ((rc = upload_columns(h, bc, h->bcpg, g_begin, g_count, h->d_bc[0] + off)))
        return rc;
    if (h->d_bc[1] != h->d_bc[0]) // Jacobi: PRESCRIBED faces are never rewritten, keep both copies alike
        rc = upload_columns(h, bc, h->bcpg, g_begin, g_count, h->d_bc[1] + off);
    return rc;
}

int mocb200_get_boundary(mocb200_sweeper *h, int plane, int g_begin, int g_count, double *bc)
{
    int rc = check_groups(h, g_begin, g_count);
    if (rc)
        return rc;
    if (plane < 0 || plane >= h->n_plane || !bc)
        return fail(h, MOCB200_ERR_INVALID, "get_boundary: bad plane %d or NULL array", plane);
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t off = (size_t)plane * h->bcpg * h->GP;
    return download_columns(h, bc, h->bcpg, g_begin, g_count, h->d_bc[h->bc_cur] + off);
}

} // extern "C"

// finalize == false: one inner sweep that leaves the FSR tally un-normalised (mocb200_sweep_partial)
static int sweep_impl(mocb200_sweeper *h, int g_begin, int g_count, int n_inner, int tally_mode, int use_qbar, bool finalize);

extern "C" {

int mocb200_sweep(mocb200_sweeper *h, int g_begin, int g_count, int n_inner, int tally_mode, int use_qbar)
{
    if (h && h->family_partial)
        return fail(h, MOCB200_ERR_STATE, "this handle sweeps a subset of the angle families: use mocb200_sweep_partial, "
                                          "sum the tallies over the ranks, then mocb200_finalize_flux");
    return sweep_impl(h, g_begin, g_count, n_inner, tally_mode, use_qbar, true);
}

int mocb200_sweep_partial(mocb200_sweeper *h, int g_begin, int g_count, int tally_mode, int use_qbar)
{
    if (h && g_count > 2)
        return fail(h, MOCB200_ERR_INVALID, "mocb200_sweep_partial: at most two groups per call (per-group sweep kernels)");
    if (h && tally_mode == MOCB200_TALLY_CORRECTIONS)
        return fail(h, MOCB200_ERR_INVALID, "mocb200_sweep_partial: the 2D3D correction factors need every angle on one handle");
    return sweep_impl(h, g_begin, g_count, 1, tally_mode, use_qbar, false);
}

int mocb200_finalize_flux(mocb200_sweeper *h, int g_begin, int g_count)
{
    int rc = check_groups(h, g_begin, g_count);
    if (rc)
        return rc;
    if (g_count > 2)
        return fail(h, MOCB200_ERR_INVALID, "mocb200_finalize_flux: at most two groups per call");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const bool cached = h->kernel == MOCB200_KERNEL_CACHED || h->kernel == MOCB200_KERNEL_CHUNK || h->kernel == MOCB200_KERNEL_RCHUNK;
    const bool rchunk = h->kernel == MOCB200_KERNEL_RCHUNK;
    const int64_t nrg = (int64_t)(h->reg_hi - h->reg_lo) * g_count;
    finalize_flux_q_kernel<<<grid_for(nrg, 256, h->sm_count), 256, 0, h->stream>>>(
        h->n_reg, h->GP, g_begin, g_count, cached ? h->d_tg : h->d_tally, h->d_xstr, h->d_vol, h->d_qbar, h->d_flux, h->reg_lo,
        h->reg_hi, cached ? 1 : 0, rchunk ? h->d_fsr_perm : nullptr, h->n_regp);
    h->stats.kernel_launches++;
    CUDA_TRY(h, cudaGetLastError());
    return MOCB200_OK;
}

int mocb200_angle_families(const mocb200_problem *prob, int32_t *n_family, int32_t *family_of_angle)
{
    if (!prob || !n_family)
        return MOCB200_ERR_INVALID;
    std::vector<int> fam;
    *n_family = angle_families(*prob, fam);
    if (family_of_angle)
        for (int i = 0; i < 2 * prob->n_ang; i++)
            family_of_angle[i] = fam[i];
    return MOCB200_OK;
}

int mocb200_device_buffer(mocb200_sweeper *h, int which, void **ptr, int64_t *count)
{
    if (!h || !ptr || !count)
        return MOCB200_ERR_INVALID;
    const bool cached = h->kernel == MOCB200_KERNEL_CACHED || h->kernel == MOCB200_KERNEL_CHUNK || h->kernel == MOCB200_KERNEL_RCHUNK;
    switch (which) {
    case MOCB200_BUF_TALLY:
        *ptr   = cached ? h->d_tg : h->d_tally;
        *count = cached ? (int64_t)h->G * std::max(h->n_reg, h->n_regp) : (int64_t)h->n_reg * h->GP;
        return MOCB200_OK;
    case MOCB200_BUF_CURRENT: *ptr = h->d_current, *count = (int64_t)h->n_surf * h->GP; return MOCB200_OK;
    case MOCB200_BUF_SURFACE_FLUX: *ptr = h->d_surfflux, *count = (int64_t)h->n_surf * h->GP; return MOCB200_OK;
    }
    return fail(h, MOCB200_ERR_INVALID, "unknown device buffer %d", which);
}

int mocb200_adopt_device_buffer(mocb200_sweeper *h, int which, void *ptr, int64_t count)
{
    void *old    = nullptr;
    int64_t need = 0;
    int rc       = mocb200_device_buffer(h, which, &old, &need);
    if (rc)
        return rc;
    if (!ptr || count < need)
        return fail(h, MOCB200_ERR_INVALID, "adopted buffer holds %lld doubles, %lld needed", (long long)count, (long long)need);
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaMemcpy(ptr, old, (size_t)need * sizeof(double), cudaMemcpyDeviceToDevice));
    const bool cached = h->kernel == MOCB200_KERNEL_CACHED || h->kernel == MOCB200_KERNEL_CHUNK || h->kernel == MOCB200_KERNEL_RCHUNK;
    switch (which) { // the handle's own allocation stays in its list and is freed with the handle
    case MOCB200_BUF_TALLY: (cached ? h->d_tg : h->d_tally) = (double *)ptr; break;
    case MOCB200_BUF_CURRENT: h->d_current = (double *)ptr; break;
    case MOCB200_BUF_SURFACE_FLUX: h->d_surfflux = (double *)ptr; break;
    }
    return MOCB200_OK;
}

} // extern "C"

static int sweep_impl(mocb200_sweeper *h, int g_begin, int g_count, int n_inner, int tally_mode, int use_qbar, bool finalize)
{
    int rc = check_groups(h, g_begin, g_count);
    if (rc)
        return rc;
    if (n_inner < 1)
        return fail(h, MOCB200_ERR_INVALID, "n_inner must be >= 1");
    if (use_qbar && n_inner > 1)
        return fail(h, MOCB200_ERR_INVALID, "use_qbar sweeps the q-bar of mocb200_set_qbar as it is: n_inner must be 1");
    if (tally_mode < MOCB200_TALLY_NONE || tally_mode > MOCB200_TALLY_CORRECTIONS)
        return fail(h, MOCB200_ERR_INVALID, "unknown tally mode %d", tally_mode);
    if (tally_mode == MOCB200_TALLY_CORRECTIONS) {
        if (h->kernel == MOCB200_KERNEL_ITEM)
            return fail(h, MOCB200_ERR_INVALID, "the item kernel has no correction-factor tally");
        if (!h->have_corr)
            return fail(h, MOCB200_ERR_STATE, "2D3D correction factors unavailable: %s", h->corr_why.c_str());
        for (int g = g_begin; g < g_begin + g_count; g++)
            if (!h->have_sn_xs[g])
                return fail(h, MOCB200_ERR_STATE, "mocb200_set_sn_xs has not been called for group %d", g);
    }
    for (int g = g_begin; g < g_begin + g_count; g++)
        if (!h->have_xs[g])
            return fail(h, MOCB200_ERR_STATE, "mocb200_set_xs has not been called for group %d", g);
    CUDA_TRY(h, cudaSetDevice(h->device));
    const bool jacobi = h->opt.boundary_update == MOCB200_BOUNDARY_JACOBI;
    const int block   = h->opt.block_threads > 0 ? std::min(512, (h->opt.block_threads + 31) & ~31) : 512;
    const int smem    = (h->exp_n + 2) * (int)sizeof(double);
    const int64_t nrg = (int64_t)(h->reg_hi - h->reg_lo) * g_count;

    const int gl          = g_count <= 2 ? 1 : 8; // group lanes per segment (track / cached kernels)
    const int n_gsets     = (g_count + gl - 1) / gl;
    const bool cached     = h->kernel == MOCB200_KERNEL_CACHED || h->kernel == MOCB200_KERNEL_CHUNK ||
                            h->kernel == MOCB200_KERNEL_RCHUNK;
    const bool group_major = cached && gl == 1;
    const bool rchunk      = h->kernel == MOCB200_KERNEL_RCHUNK && gl == 1; // packed-batch layout of the cache

    if (cached && gl != 1 && h->cache_slots < h->G)
        return fail(h, MOCB200_ERR_STATE, "group-batched sweeps need the attenuation cache of all groups (%d of %d fit)",
                    h->cache_slots, h->G);
    if (cached && g_count > h->cache_slots)
        return fail(h, MOCB200_ERR_STATE, "%d groups per call but the attenuation cache holds %d", g_count, h->cache_slots);
    const bool sliding = cached && h->cache_slots < h->G; // the cache holds the groups of this call only
    if (cached) {
        // (re)build the attenuation cache for groups whose cross sections changed
        if (h->cache_layout != gl) {
            for (auto &tl : h->tlists) {
                if (tl.d_cache) {
                    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
                    CUDA_TRY(h, cudaFree(tl.d_cache));
                    h->device_bytes -= tl.cache_positions(h->kernel == MOCB200_KERNEL_RCHUNK && h->cache_layout == 1) * tl.np *
                                       tl.n_planes * (int64_t)(h->cache_layout == 1 ? h->cache_slots : h->GP) * 8;
                    tl.d_cache = nullptr;
                }
                const size_t bytes = (size_t)tl.cache_positions(rchunk) * tl.np * tl.n_planes * (gl == 1 ? h->cache_slots : h->GP) * 8;
                CUDA_TRY(h, cudaMalloc((void **)&tl.d_cache, bytes));
                h->device_bytes += (int64_t)bytes;
            }
            h->cache_layout = gl;
            h->cache_valid.assign(h->G, false);
            h->stats.device_bytes = h->device_bytes;
        }
        if (sliding && (h->cache_g0 != g_begin || h->cache_gn != g_count)) {
            h->cache_valid.assign(h->G, false); // other groups are resident: this call's replace them
            h->cache_g0 = g_begin, h->cache_gn = g_count;
        }
        bool dirty = false;
        for (int g = g_begin; g < g_begin + g_count; g++)
            dirty = dirty || !h->cache_valid[g];
        if (dirty) {
            for (const auto &tl : h->tlists) {
                if (rchunk) {
                    RcCacheArgs c{};
                    c.chunk_trk = tl.rc.d_chunk_trk, c.tracks = tl.d_cunits, c.len_begin = tl.d_len_begin;
                    c.planes = tl.d_planes, c.n_planes = tl.n_planes, c.lmax = tl.rc.LMAX, c.n_slots = tl.rc.n_slots;
                    c.seg_len = h->d_pseg_len, c.seg_fsr = h->d_pseg_fsr, c.ang_rsintheta = h->d_rsin;
                    c.plane_first_reg = h->d_plane_first_reg, c.xstr = h->d_xstr;
                    c.g_begin = g_begin, c.g_count = g_count, c.cache_groups = h->cache_slots;
                    c.cache_g0 = sliding ? g_begin : 0, c.GP = h->GP, c.cache = tl.d_cache;
                    c.exp_table = h->d_exp, c.exp_n = h->exp_n, c.exp_min = h->exp_min, c.exp_max = h->exp_max;
                    const int grid = (int)std::max<int64_t>(
                        1, std::min<int64_t>((tl.rc.n_slots * tl.n_planes + 511) / 512, (int64_t)h->sm_count));
                    pick_rc_cache_kernel(tl.np)<<<grid, 512, smem, h->stream>>>(c);
                    h->stats.kernel_launches++;
                    continue;
                }
                CacheArgs c{};
                c.units = tl.d_units, c.n_units = tl.n_units, c.bundles = h->d_tbundles, c.len_begin = tl.d_len_begin;
                c.planes = tl.d_planes, c.n_planes = tl.n_planes;
                c.seg_len = h->d_pseg_len, c.seg_fsr = h->d_pseg_fsr, c.ang_rsintheta = h->d_rsin;
                c.plane_first_reg = h->d_plane_first_reg, c.xstr = h->d_xstr;
                c.g_begin = g_begin, c.g_count = g_count, c.cache_groups = h->cache_slots;
                c.cache_g0 = sliding ? g_begin : 0;
                c.GP = h->GP, c.np = tl.np, c.group_major = gl == 1 ? 1 : 0;
                c.cache = tl.d_cache, c.list_pseg = tl.pseg;
                c.exp_table = h->d_exp, c.exp_n = h->exp_n, c.exp_min = h->exp_min, c.exp_max = h->exp_max;
                const int64_t warps = (int64_t)tl.n_units * tl.n_planes;
                const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((warps + 15) / 16, h->sm_count));
                pick_cache_kernel(tl.np)<<<grid, 512, smem, h->stream>>>(c);
                h->stats.kernel_launches++;
            }
            for (int g = g_begin; g < g_begin + g_count; g++)
                h->cache_valid[g] = true;
        }
    }

    if (tally_mode == MOCB200_TALLY_CORRECTIONS) {
        // per-angle, per-direction sums of the corrections sweep and the resulting factors
        const size_t ngl = gl == 1 ? (size_t)g_count : (size_t)h->GP;
        const size_t nd  = (size_t)h->n_reg * 2 * h->n_ang * ngl;
        const size_t ns  = (size_t)h->n_plane * h->n_ang * h->n_surf_plane * 2 * ngl;
        if (nd > h->dsum_elems || ns > h->ssum_elems || (size_t)g_count > h->corr_groups) {
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            for (double **q : {&h->d_dsum, &h->d_ssum, &h->d_alpha, &h->d_beta})
                if (*q) {
                    CUDA_TRY(h, cudaFree(*q));
                    *q = nullptr;
                }
            const size_t nc = (size_t)g_count * 2 * h->n_ang * h->n_plane * h->n_cell_plane;
            CUDA_TRY(h, cudaMalloc((void **)&h->d_dsum, nd * sizeof(double)));
            CUDA_TRY(h, cudaMalloc((void **)&h->d_ssum, ns * sizeof(double)));
            CUDA_TRY(h, cudaMalloc((void **)&h->d_alpha, 2 * nc * sizeof(double)));
            CUDA_TRY(h, cudaMalloc((void **)&h->d_beta, nc * sizeof(double)));
            h->dsum_elems = nd, h->ssum_elems = ns, h->corr_groups = (size_t)g_count;
        }
    }

    // ---- persistent path: every plain inner of this call in one cooperative launch (moc_rchunk_kernel.cuh) ----
    int inner0 = 0;
    if (rchunk && !use_qbar && h->n_counters <= 256 && h->persist_mode > 0 && finalize) {
        const TrackList *pl[2] = {nullptr, nullptr};
        int n_in_phase[2]      = {0, 0};
        for (const auto &tl : h->tlists)
            if (tl.phase == 0 || tl.phase == 1)
                pl[tl.phase] = &tl, n_in_phase[tl.phase]++;
        const int n_phases = jacobi ? 1 : 2;
        bool ok            = n_in_phase[0] == 1 && n_in_phase[1] == (jacobi ? 0 : 1);
        if (ok && !jacobi)
            ok = pl[0]->rc.P == pl[1]->rc.P && pl[0]->rc.LMAX == pl[1]->rc.LMAX && pl[0]->rc.NW == pl[1]->rc.NW &&
                 pl[0]->rc.TEAMS == pl[1]->rc.TEAMS;
        const int m = tally_mode != MOCB200_TALLY_NONE ? n_inner - 1 : n_inner;
        RcPersistFn pfn = nullptr;
        if (ok && m >= 1)
            pfn = pick_rc_persist_kernel(pl[0]->rc.P, RcConfig{pl[0]->rc.LMAX, pl[0]->rc.NW, pl[0]->rc.TEAMS});
        if (pfn) {
            const RcList &rc0 = pl[0]->rc;
            const size_t dsm  = (size_t)rc0.TEAMS * rc_team_bytes(rc0.P, rc0.LMAX, rc0.NW);
            if (!h->persist_attr_set[rc0.P]) {
                CUDA_TRY(h, cudaFuncSetAttribute((const void *)pfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
                h->persist_attr_set[rc0.P] = true;
            }
            const int ss_grid = grid_for((int64_t)h->n_reg * g_count, 256, h->sm_count);
            self_scatter_q_kernel<<<ss_grid, 256, 0, h->stream>>>(
                h->n_reg, h->GP, g_begin, g_count, h->d_src, h->d_flux, h->d_xs_self, h->d_xstr_src, h->d_xstr, h->d_qbar,
                h->d_qg, nullptr, h->d_tg, 1, 1, h->d_counters, h->n_counters, h->d_fsr_perm, h->n_regp);
            h->stats.kernel_launches++;
            CUDA_TRY(h, cudaMemsetAsync(h->d_gridbar, 0, sizeof(unsigned int), h->stream));
            RcPersistArgs pa{};
            int64_t max_items = 0;
            for (int ph = 0; ph < n_phases; ph++) {
                const TrackList &tl = *pl[ph];
                const RcList &rc    = tl.rc;
                RcArgs &c           = pa.ph[ph];
                c.units = rc.d_units, c.n_units = rc.n_units, c.pinfo = rc.d_pinfo, c.n_planes = tl.n_planes;
                c.plane_first_reg = h->d_plane_first_reg, c.seg_fsr = h->d_pseg_fsr, c.n_regp = h->n_regp;
                c.chunk_trk = rc.d_chunk_trk;
                c.tracks = tl.d_cunits, c.batch_hdr = rc.d_batch_hdr, c.cache = tl.d_cache, c.n_slots = rc.n_slots;
                c.cache_groups = h->cache_slots, c.cache_g0 = sliding ? g_begin : 0;
                c.wt_v_st = h->d_wt, c.n_ang = h->n_ang, c.bc_per_group = h->bcpg;
                c.g_begin = g_begin, c.g_count = g_count, c.GP = h->GP, c.n_reg = h->n_reg;
                c.q = h->d_qg, c.tally = h->d_tg;
                c.scratch = h->d_scratch, c.scratch_per_team = h->rc_scratch_per_team;
                c.interleave = h->rc_interleave;
                max_items = std::max(max_items, (int64_t)rc.n_units * tl.n_planes * g_count);
            }
            pa.n_phases = n_phases, pa.n_inner = m, pa.finalize_last = m < n_inner ? 1 : 0;
            pa.prefetch = h->persist_mode - 1;
            pa.bc[0] = h->d_bc[h->bc_cur], pa.bc[1] = jacobi ? h->d_bc[1 - h->bc_cur] : h->d_bc[h->bc_cur];
            RcFinalizeArgs &f = pa.fin;
            f.n_reg = h->n_reg, f.GP = h->GP, f.g_begin = g_begin, f.g_count = g_count, f.reg_lo = h->reg_lo;
            f.reg_hi = h->reg_hi, f.n_regp = h->n_regp, f.tally = h->d_tg, f.xstr = h->d_xstr, f.vol = h->d_vol;
            f.qbar = h->d_qbar, f.flux = h->d_flux, f.src = h->d_src, f.xs_self = h->d_xs_self;
            f.xstr_src = h->d_xstr_src, f.q_out = h->d_qg, f.perm = h->d_fsr_perm;
            pa.barrier = h->d_gridbar;
            const int pgrid =
                (int)std::max<int64_t>(1, std::min<int64_t>((max_items + rc0.TEAMS - 1) / rc0.TEAMS, h->sm_count));
            const bool whole = m == n_inner;
            if (whole)
                CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
            if (h->timing) {
                if (h->ev_used == h->ev_pool.size()) {
                    cudaEvent_t a = nullptr, b = nullptr;
                    CUDA_TRY(h, cudaEventCreate(&a));
                    CUDA_TRY(h, cudaEventCreate(&b));
                    h->ev_pool.emplace_back(a, b);
                    h->ev_inners.push_back(1);
                }
                h->ev_inners[h->ev_used] = m;
                CUDA_TRY(h, cudaEventRecord(h->ev_pool[h->ev_used].first, h->stream));
            }
            void *kargs[] = {(void *)&pa};
            CUDA_TRY(h, cudaLaunchCooperativeKernel((const void *)pfn, dim3(pgrid), dim3(32 * rc0.NW * rc0.TEAMS), kargs, dsm,
                                                    h->stream));
            h->stats.kernel_launches++;
            h->stats.sweep_launches++;
            if (h->timing)
                CUDA_TRY(h, cudaEventRecord(h->ev_pool[h->ev_used++].second, h->stream));
            if (jacobi && (m & 1))
                h->bc_cur = 1 - h->bc_cur;
            if (whole) {
                CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
                h->ev_valid = true;
                finalize_flux_q_kernel<<<grid_for(nrg, 256, h->sm_count), 256, 0, h->stream>>>(
                    h->n_reg, h->GP, g_begin, g_count, h->d_tg, h->d_xstr, h->d_vol, h->d_qbar, h->d_flux, h->reg_lo,
                    h->reg_hi, 1, h->d_fsr_perm, h->n_regp);
                h->stats.kernel_launches++;
                CUDA_TRY(h, cudaGetLastError());
                return MOCB200_OK;
            }
            inner0 = m;
        }
    }

    for (int inner = inner0; inner < n_inner; inner++) {
        const bool last = inner == n_inner - 1;
        const int tally = last ? tally_mode : MOCB200_TALLY_NONE;
        if (tally == MOCB200_TALLY_CORRECTIONS) {
            CUDA_TRY(h, cudaMemsetAsync(h->d_dsum, 0, h->dsum_elems * sizeof(double), h->stream));
            CUDA_TRY(h, cudaMemsetAsync(h->d_ssum, 0, h->ssum_elems * sizeof(double), h->stream));
        }
        // q-bar and tally reset (whole FSR range: cheap, keeps indexing simple)
        const int ss_grid = grid_for((int64_t)h->n_reg * g_count, 256, h->sm_count);
        // per-group sweeps fuse the flux update of inner i with the q-bar of inner i + 1 (see the end of the loop)
        const bool fuse_q = group_major && !use_qbar && h->n_counters <= 256;
        if (h->kernel == MOCB200_KERNEL_ITEM) {
            self_scatter_kernel<<<ss_grid, 256, 0, h->stream>>>(
                h->n_reg, h->GP, g_begin, g_count, h->d_src, h->d_flux, h->d_xs_self, h->d_xstr_src, h->d_qbar,
                h->d_tally, use_qbar ? 0 : 1);
            h->stats.kernel_launches++;
        } else if (inner == 0 || !fuse_q) {
            self_scatter_q_kernel<<<ss_grid, 256, 0, h->stream>>>(
                h->n_reg, h->GP, g_begin, g_count, h->d_src, h->d_flux, h->d_xs_self, h->d_xstr_src, h->d_xstr,
                h->d_qbar, group_major ? h->d_qg : h->d_qbar, cached ? nullptr : h->d_xq,
                group_major ? h->d_tg : h->d_tally, group_major ? 1 : 0, use_qbar ? 0 : 1, h->d_counters,
                h->n_counters <= 256 ? h->n_counters : 0, rchunk ? h->d_fsr_perm : nullptr, h->n_regp);
            h->stats.kernel_launches++;
        }
        if (tally != MOCB200_TALLY_NONE) {
            zero_groups_kernel<<<grid_for((int64_t)h->n_surf * g_count, 256, h->sm_count), 256, 0, h->stream>>>(
                h->n_surf, h->GP, g_begin, g_count, h->d_current);
            zero_groups_kernel<<<grid_for((int64_t)h->n_surf * g_count, 256, h->sm_count), 256, 0, h->stream>>>(
                h->n_surf, h->GP, g_begin, g_count, h->d_surfflux);
            h->stats.kernel_launches += 2;
        }
        if (h->kernel == MOCB200_KERNEL_ITEM || h->n_counters > 256)
            CUDA_TRY(h, cudaMemsetAsync(h->d_counters, 0, sizeof(uint32_t) * h->n_counters, h->stream));
        if (last)
            CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
        if (h->timing) {
            if (h->ev_used == h->ev_pool.size()) {
                cudaEvent_t a = nullptr, b = nullptr;
                CUDA_TRY(h, cudaEventCreate(&a));
                CUDA_TRY(h, cudaEventCreate(&b));
                h->ev_pool.emplace_back(a, b);
                h->ev_inners.push_back(1);
            }
            h->ev_inners[h->ev_used] = 1;
            CUDA_TRY(h, cudaEventRecord(h->ev_pool[h->ev_used].first, h->stream));
        }
        const double *bc_in = h->d_bc[h->bc_cur];
        double *bc_out      = jacobi ? h->d_bc[1 - h->bc_cur] : h->d_bc[h->bc_cur];
        for (int phase = 0; phase < (jacobi ? 1 : 2); phase++) {
            if (h->kernel == MOCB200_KERNEL_ITEM) {
                for (size_t il = 0; il < h->lists.size(); il++) {
                    const WorkList &wl = h->lists[il];
                    if (wl.phase != phase)
                        continue;
                    SweepArgs a{};
                    a.items = wl.d_items, a.n_items = wl.n_items, a.counter = h->d_counters + il;
                    a.bundles = h->d_bundles, a.planes = wl.d_planes, a.n_planes = wl.n_planes;
                    a.seg_len = h->d_seg_len, a.seg_fsr = h->d_seg_fsr, a.cross = h->d_cross;
                    a.ang_rsintheta = h->d_rsin, a.wt_v_st = h->d_wt, a.cur_w = h->d_curw, a.flx_w = h->d_flxw;
                    a.bc_offset = h->d_bc_offset, a.bc_size_x = h->d_bc_size_x;
                    a.bc_dst_off = h->d_bc_dst_off, a.bc_dst_kind = h->d_bc_dst_kind;
                    a.plane_first_reg = h->d_plane_first_reg, a.plane_surf_offset = h->d_plane_surf_offset;
                    a.n_ang = h->n_ang, a.bc_per_group = h->bcpg;
                    a.g_begin = g_begin, a.g_count = g_count, a.GP = h->GP;
                    a.xstr = h->d_xstr, a.qbar = h->d_qbar, a.tally = h->d_tally;
                    a.bc_in = bc_in, a.bc_out = bc_out;
                    a.current = h->d_current, a.surface_flux = h->d_surfflux;
                    a.exp_table = h->d_exp, a.exp_n = h->exp_n, a.exp_min = h->exp_min, a.exp_max = h->exp_max;
                    const int64_t threads = (int64_t)wl.n_items * wl.n_planes * g_count;
                    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((threads + block - 1) / block,
                                                                                (int64_t)h->sm_count * 2));
                    pick_kernel(wl.np, tally)<<<grid, block, smem, h->stream>>>(a);
                    h->stats.kernel_launches++;
                    h->stats.sweep_launches++;
                }
                continue;
            }
            for (size_t il = 0; il < h->tlists.size(); il++) {
                const TrackList &tl = h->tlists[il];
                if (tl.phase != phase)
                    continue;
                const int64_t warps = (int64_t)tl.n_units * tl.n_planes * n_gsets;
                const int grid      = (int)std::max<int64_t>(1, std::min<int64_t>((warps + 15) / 16, h->track_grid));
                uint32_t *counter   = h->d_counters + h->lists.size() + il;
                WarpArgs a{};
                a.units = tl.d_units, a.n_units = tl.n_units, a.counter = counter;
                a.bundles = h->d_tbundles, a.planes = tl.d_planes, a.n_planes = tl.n_planes;
                a.seg_len = h->d_pseg_len, a.seg_fsr = h->d_pseg_fsr, a.xptr = h->d_xptr, a.cross = h->d_xcross;
                a.ang_rsintheta = h->d_rsin, a.wt_v_st = h->d_wt, a.cur_w = h->d_curw, a.flx_w = h->d_flxw;
                a.bc_offset = h->d_bc_offset, a.bc_size_x = h->d_bc_size_x;
                a.bc_dst_off = h->d_bc_dst_off, a.bc_dst_kind = h->d_bc_dst_kind;
                a.plane_first_reg = h->d_plane_first_reg, a.plane_surf_offset = h->d_plane_surf_offset;
                a.n_ang = h->n_ang, a.bc_per_group = h->bcpg, a.n_plane_total = h->n_plane;
                a.n_surf_plane = h->n_surf_plane;
                a.g_begin = g_begin, a.g_count = g_count, a.GP = h->GP, a.n_gsets = n_gsets, a.n_reg = h->n_reg;
                a.xq = h->d_xq, a.q = group_major ? h->d_qg : h->d_qbar;
                a.tally = group_major ? h->d_tg : h->d_tally;
                a.bc_in = bc_in, a.bc_out = bc_out;
                a.current = h->d_current, a.surface_flux = h->d_surfflux;
                a.dsum = h->d_dsum, a.ssum = h->d_ssum;
                a.scratch = h->d_scratch, a.scratch_per_warp = h->scratch_per_warp;
                a.cache = tl.d_cache, a.list_pseg = tl.pseg, a.cache_groups = h->cache_slots;
                a.cache_g0 = sliding ? g_begin : 0;
                a.exp_table = h->d_exp, a.exp_n = h->exp_n, a.exp_min = h->exp_min, a.exp_max = h->exp_max;
                if (rchunk) {
                    const RcList &rc = tl.rc;
                    const RcConfig cfg = rc_tally_config(rc.P, tally, RcConfig{rc.LMAX, rc.NW, rc.TEAMS});
                    RcFn fn = pick_rc_kernel(rc.P, tally, cfg);
                    if (!fn)
                        return fail(h, MOCB200_ERR_INVALID, "register-chunk kernel: no instantiation for P %d tally %d (%d,%d,%d)",
                                    rc.P, tally, rc.LMAX, rc.NW, rc.TEAMS);
                    RcArgs c{};
                    c.units = rc.d_units, c.n_units = rc.n_units, c.pinfo = rc.d_pinfo, c.n_planes = tl.n_planes;
                    c.plane_first_reg = h->d_plane_first_reg, c.seg_fsr = h->d_pseg_fsr, c.n_regp = h->n_regp;
                    c.chunk_trk = rc.d_chunk_trk;
                    c.tracks = tl.d_cunits, c.batch_hdr = rc.d_batch_hdr, c.cache = tl.d_cache, c.n_slots = rc.n_slots;
                    c.cache_groups = h->cache_slots, c.cache_g0 = sliding ? g_begin : 0;
                    c.wt_v_st = h->d_wt, c.n_ang = h->n_ang, c.bc_per_group = h->bcpg;
                    c.g_begin = g_begin, c.g_count = g_count, c.GP = h->GP, c.n_reg = h->n_reg;
                    c.q = h->d_qg, c.tally = h->d_tg, c.bc_in = bc_in, c.bc_out = bc_out;
                    c.scratch = h->d_scratch, c.scratch_per_team = h->rc_scratch_per_team;
                    c.interleave = h->rc_interleave;
                    c.cross = h->d_xcross, c.cur_w = h->d_curw, c.flx_w = h->d_flxw;
                    c.plane_surf_offset = h->d_plane_surf_offset, c.current = h->d_current, c.surface_flux = h->d_surfflux;
                    c.dsum = h->d_dsum, c.ssum = h->d_ssum, c.n_surf_plane = h->n_surf_plane, c.n_plane_total = h->n_plane;
                    const int64_t items = (int64_t)rc.n_units * tl.n_planes * g_count;
                    const int rgrid = (int)std::max<int64_t>(1, std::min<int64_t>((items + cfg.TEAMS - 1) / cfg.TEAMS, h->sm_count));
                    fn<<<rgrid, 32 * rc.NW * cfg.TEAMS, cfg.TEAMS * rc_team_bytes(rc.P, rc.LMAX, rc.NW), h->stream>>>(c);
                } else if (h->kernel == MOCB200_KERNEL_CHUNK && gl == 1) {
                    int caps = 0, nw = 1, teams = 1;
                    const bool tl_tally = tally != MOCB200_TALLY_NONE;
                    chunk_geometry(tl.nseg_desc, tl.np, h->opt.chunk_cap, tl_tally, &caps, &nw, &teams);
                    if (h->opt.chunk_cap < 0) // test hook: negative cap = that cap with two-warp teams
                        chunk_geometry(tl.nseg_desc, tl.np, -h->opt.chunk_cap, tl_tally, &caps, &nw, &teams), nw = 2,
                            teams = std::min(teams, chunk_max_warps(2) / 2);
                    const int cgrid =
                        (int)std::max<int64_t>(1, std::min<int64_t>((warps + teams - 1) / teams, h->track_grid));
                    ChunkArgs c{};
                    c.units = tl.d_cunits, c.n_units = tl.n_units, c.counter = counter;
                    c.pinfo = tl.d_pinfo, c.n_planes = tl.n_planes, c.seg_fsr = h->d_pseg_fsr, c.wt_v_st = h->d_wt;
                    c.n_ang = h->n_ang, c.bc_per_group = h->bcpg;
                    c.g_begin = g_begin, c.g_count = g_count, c.GP = h->GP, c.n_reg = h->n_reg;
                    c.q = h->d_qg, c.tally = h->d_tg, c.bc_in = bc_in, c.bc_out = bc_out;
                    c.scratch = h->d_scratch, c.scratch_per_warp = h->scratch_per_warp;
                    c.cache = tl.d_cache, c.list_pseg = tl.pseg, c.cache_groups = h->cache_slots, c.caps = caps;
                    c.cache_g0 = sliding ? g_begin : 0;
                    static const char *ex_mode = getenv("MOCB200_CHUNK_EX"); // tuning hook
                    c.ex_mode = ex_mode && ex_mode[0] == '1' ? 1 : 0;
                    c.xptr = h->d_xptr, c.cross = h->d_xcross, c.cur_w = h->d_curw, c.flx_w = h->d_flxw;
                    c.plane_surf_offset = h->d_plane_surf_offset, c.current = h->d_current, c.surface_flux = h->d_surfflux;
                    c.dsum = h->d_dsum, c.ssum = h->d_ssum, c.n_surf_plane = h->n_surf_plane, c.n_plane_total = h->n_plane;
                    pick_chunk_kernel(tl.np, nw, tally)<<<cgrid, 32 * nw * teams, teams * chunk_warp_bytes(caps, tl.np, tl_tally),
                                                   h->stream>>>(c);
                } else {
                    pick_warp_kernel(gl, tl.np, tally, cached)<<<grid, kWarpBlock, cached ? 0 : track_smem_bytes(h),
                                                                 h->stream>>>(a);
                }
                h->stats.kernel_launches++;
                h->stats.sweep_launches++;
            }
        }
        if (h->timing)
            CUDA_TRY(h, cudaEventRecord(h->ev_pool[h->ev_used++].second, h->stream));
        if (last) {
            CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
            h->ev_valid = true;
        }
        if (tally == MOCB200_TALLY_CORRECTIONS) {
            CorrArgs c{};
            c.planes = h->d_all_planes, c.n_planes = h->plane_end - h->plane_begin, c.n_ang = h->n_ang;
            c.n_cell_plane = h->n_cell_plane, c.n_surf_plane = h->n_surf_plane, c.n_plane_total = h->n_plane;
            c.n_geom = h->n_geom, c.GP = h->GP, c.n_reg = h->n_reg, c.gl = gl;
            c.g_begin = g_begin, c.g_count = g_count;
            c.plane_unique = h->d_plane_unique, c.plane_first_reg = h->d_plane_first_reg;
            c.plane_cell_offset = h->d_plane_cell_offset, c.uniq_reg_begin = h->d_uniq_reg_begin;
            c.cell_fsr_begin = h->d_cell_fsr_begin, c.cell_fsr = h->d_cell_fsr, c.geom_len = h->d_geom_len;
            c.ang_geom = h->d_ang_geom, c.ang_rsintheta = h->d_rsin;
            c.ang_area_x = h->d_area_x, c.ang_area_y = h->d_area_y, c.ang_ox = h->d_ox;
            c.cell_dx = h->d_cell_dx, c.cell_dy = h->d_cell_dy, c.coarse_surf = h->d_coarse_surf;
            c.xstr = h->d_xstr, c.xstr_true = h->d_xstr_src, c.qbar = h->d_qbar, c.sn_xs = h->d_sn_xs;
            c.dsum = h->d_dsum, c.ssum = h->d_ssum, c.alpha = h->d_alpha, c.beta = h->d_beta;
            const int64_t n = (int64_t)c.n_planes * c.n_ang * c.n_cell_plane * 2 * g_count;
            corrections_kernel<<<grid_for(n, 128, h->sm_count), 128, 0, h->stream>>>(c);
            h->stats.kernel_launches++;
            h->corr_g_begin = g_begin, h->corr_g_count = g_count;
        }
        if (jacobi)
            h->bc_cur = 1 - h->bc_cur;
        if (!finalize) { // mocb200_sweep_partial: the caller sums the tally over the ranks, then mocb200_finalize_flux
            CUDA_TRY(h, cudaGetLastError());
            continue;
        }
        if (fuse_q && !last)
            finalize_next_q_kernel<<<grid_for(nrg, 256, h->sm_count), 256, 0, h->stream>>>(
                h->n_reg, h->GP, g_begin, g_count, h->d_tg, h->d_xstr, h->d_vol, h->d_qbar, h->d_flux, h->reg_lo,
                h->reg_hi, h->d_src, h->d_xs_self, h->d_xstr_src, h->d_qg, h->d_counters, h->n_counters,
                rchunk ? h->d_fsr_perm : nullptr, h->n_regp);
        else
            finalize_flux_q_kernel<<<grid_for(nrg, 256, h->sm_count), 256, 0, h->stream>>>(
                h->n_reg, h->GP, g_begin, g_count, group_major ? h->d_tg : h->d_tally, h->d_xstr, h->d_vol, h->d_qbar,
                h->d_flux, h->reg_lo, h->reg_hi, group_major ? 1 : 0, rchunk ? h->d_fsr_perm : nullptr, h->n_regp);
        h->stats.kernel_launches++;
        CUDA_TRY(h, cudaGetLastError());
    }
    return MOCB200_OK;
}

extern "C" {

int mocb200_set_sweep_inputs(mocb200_sweeper *h, int group, const double *source, const double *flux,
                             const double *const *boundary)
{
    int rc = check_groups(h, group, 1);
    if (rc)
        return rc;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->ev_in_pending) { // the previous fused upload must have left the pinned buffer
        CUDA_TRY(h, cudaEventSynchronize(h->ev_in));
        h->ev_in_pending = false;
    }
    // only the FSR range of this handle's macroplanes travels (the host arrays keep their global indexing)
    const size_t nr = (size_t)(h->reg_hi - h->reg_lo), nb = (size_t)h->bcpg;
    size_t off = 0;
    bool staged = false;
    // pageable arrays are packed into the pinned staging buffer (one copy for all of them, flushed at the
    // end); page-locked arrays are copied straight from where they are (no host memcpy)
    size_t run_begin = 0;
    auto flush_staged = [&](size_t end) -> int {
        if (end > run_begin) {
            CUDA_TRY(h, cudaMemcpyAsync(h->d_in + run_begin, h->h_in + run_begin, (end - run_begin) * sizeof(double),
                                        cudaMemcpyHostToDevice, h->stream));
            staged = true;
        }
        return MOCB200_OK;
    };
    auto put = [&](const double *src, size_t n) -> int {
        if (is_pinned(src)) {
            int rc2 = flush_staged(off);
            if (rc2)
                return rc2;
            CUDA_TRY(h, cudaMemcpyAsync(h->d_in + off, src, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            run_begin = off + n;
        } else {
            std::memcpy(h->h_in + off, src, n * sizeof(double));
        }
        off += n;
        return MOCB200_OK;
    };
    const size_t o_src = off;
    if (source && (rc = put(source + h->reg_lo, nr)))
        return rc;
    const size_t o_flux = off;
    if (flux && (rc = put(flux + h->reg_lo, nr)))
        return rc;
    const size_t o_bc = off;
    std::vector<int> bc_planes;
    for (int ip = h->plane_begin; boundary && ip < h->plane_end; ip++) {
        if (!boundary[ip])
            continue;
        if ((rc = put(boundary[ip], nb)))
            return rc;
        bc_planes.push_back(ip);
    }
    if (off == 0)
        return MOCB200_OK;
    if ((rc = flush_staged(off)))
        return rc;
    if (staged) {
        CUDA_TRY(h, cudaEventRecord(h->ev_in, h->stream));
        h->ev_in_pending = true;
    }
    auto scatter = [&](const double *cols, int64_t n, double *dst) {
        scatter_columns_kernel<<<grid_for(n, 256, h->sm_count), 256, 0, h->stream>>>(n, h->GP, group, 1, cols, dst);
        h->stats.kernel_launches++;
    };
    if (source)
        scatter(h->d_in + o_src, (int64_t)nr, h->d_src + (size_t)h->reg_lo * h->GP);
    if (flux)
        scatter(h->d_in + o_flux, (int64_t)nr, h->d_flux + (size_t)h->reg_lo * h->GP);
    for (size_t i = 0; i < bc_planes.size(); i++) {
        const size_t poff = (size_t)bc_planes[i] * h->bcpg * h->GP;
        scatter(h->d_in + o_bc + i * nb, h->bcpg, h->d_bc[0] + poff);
        if (h->d_bc[1] != h->d_bc[0]) // Jacobi: PRESCRIBED faces are never rewritten, keep both copies alike
            scatter(h->d_in + o_bc + i * nb, h->bcpg, h->d_bc[1] + poff);
    }
    CUDA_TRY(h, cudaGetLastError());
    return MOCB200_OK;
}

int mocb200_get_sweep_results(mocb200_sweeper *h, int group, double *flux, double *const *boundary, double *current,
                              double *surface_flux)
{
    int rc = check_groups(h, group, 1);
    if (rc)
        return rc;
    if ((current == nullptr) != (surface_flux == nullptr))
        return fail(h, MOCB200_ERR_INVALID, "get_sweep_results: current and surface_flux go together");
    CUDA_TRY(h, cudaSetDevice(h->device));
    // coarse surfaces of this handle's macroplanes (the last plane's top faces ride along: never tallied radially)
    const size_t s_lo = (size_t)h->plane_begin * h->n_surf_plane;
    const size_t s_hi = h->plane_end == h->n_plane ? (size_t)h->n_surf : (size_t)h->plane_end * h->n_surf_plane;
    const size_t nr = (size_t)(h->reg_hi - h->reg_lo), nb = (size_t)h->bcpg, ns = s_hi - s_lo;
    auto gather = [&](const double *src, int64_t n, double *cols) {
        gather_columns_kernel<<<grid_for(n, 256, h->sm_count), 256, 0, h->stream>>>(n, h->GP, group, 1, src, cols);
        h->stats.kernel_launches++;
    };
    size_t off = 0;
    const size_t o_flux = off;
    if (flux) {
        gather(h->d_flux + (size_t)h->reg_lo * h->GP, (int64_t)nr, h->d_out + off);
        off += nr;
    }
    const size_t o_bc = off;
    std::vector<int> bc_planes;
    for (int ip = h->plane_begin; boundary && ip < h->plane_end; ip++) {
        if (!boundary[ip])
            continue;
        gather(h->d_bc[h->bc_cur] + (size_t)ip * h->bcpg * h->GP, h->bcpg, h->d_out + off);
        off += nb;
        bc_planes.push_back(ip);
    }
    const size_t o_cur = off;
    if (current) {
        gather(h->d_current + s_lo * h->GP, (int64_t)ns, h->d_out + off);
        gather(h->d_surfflux + s_lo * h->GP, (int64_t)ns, h->d_out + off + ns);
        off += 2 * ns;
    }
    if (off == 0)
        return MOCB200_OK;
    CUDA_TRY(h, cudaGetLastError());
    // page-locked targets receive their part directly; the rest goes through the pinned staging buffer
    struct Piece {
        double *dst;
        size_t off, n;
        bool direct;
    };
    std::vector<Piece> pieces;
    if (flux)
        pieces.push_back({flux + h->reg_lo, o_flux, nr, is_pinned(flux)});
    for (size_t i = 0; i < bc_planes.size(); i++)
        pieces.push_back({boundary[bc_planes[i]], o_bc + i * nb, nb, is_pinned(boundary[bc_planes[i]])});
    if (current) {
        pieces.push_back({current + s_lo, o_cur, ns, is_pinned(current)});
        pieces.push_back({surface_flux + s_lo, o_cur + ns, ns, is_pinned(surface_flux)});
    }
    bool any_staged = false;
    for (const Piece &pc : pieces) {
        if (pc.direct)
            CUDA_TRY(h, cudaMemcpyAsync(pc.dst, h->d_out + pc.off, pc.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        else
            any_staged = true;
    }
    if (any_staged) // one copy for every staged piece (the direct ones are copied twice: simpler than splitting runs)
        CUDA_TRY(h, cudaMemcpyAsync(h->h_out, h->d_out, off * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    for (const Piece &pc : pieces)
        if (!pc.direct)
            std::memcpy(pc.dst, h->h_out + pc.off, pc.n * sizeof(double));
    return MOCB200_OK;
}

int mocb200_pack_results_device(mocb200_sweeper *h, int group, double *dst_device, size_t capacity, size_t *count)
{
    int rc = check_groups(h, group, 1);
    if (rc)
        return rc;
    if (!count)
        return fail(h, MOCB200_ERR_INVALID, "pack_results_device: count is NULL");
    const size_t s_lo = (size_t)h->plane_begin * h->n_surf_plane;
    const size_t s_hi = h->plane_end == h->n_plane ? (size_t)h->n_surf : (size_t)h->plane_end * h->n_surf_plane;
    const size_t nr = (size_t)(h->reg_hi - h->reg_lo), ns = s_hi - s_lo;
    *count = nr + 2 * ns;
    if (!dst_device)
        return MOCB200_OK; // size query
    if (capacity < *count)
        return fail(h, MOCB200_ERR_INVALID, "pack_results_device: buffer holds %zu doubles, %zu needed", capacity, *count);
    CUDA_TRY(h, cudaSetDevice(h->device));
    auto gather = [&](const double *src, int64_t n, double *cols) {
        gather_columns_kernel<<<grid_for(n, 256, h->sm_count), 256, 0, h->stream>>>(n, h->GP, group, 1, src, cols);
        h->stats.kernel_launches++;
    };
    gather(h->d_flux + (size_t)h->reg_lo * h->GP, (int64_t)nr, dst_device);
    gather(h->d_current + s_lo * h->GP, (int64_t)ns, dst_device + nr);
    gather(h->d_surfflux + s_lo * h->GP, (int64_t)ns, dst_device + nr + ns);
    CUDA_TRY(h, cudaGetLastError());
    return MOCB200_OK;
}

int mocb200_get_coarse(mocb200_sweeper *h, int group, double *current, double *surface_flux)
{
    int rc = check_groups(h, group, 1);
    if (rc)
        return rc;
    if (!current || !surface_flux)
        return fail(h, MOCB200_ERR_INVALID, "get_coarse: NULL array");
    CUDA_TRY(h, cudaSetDevice(h->device));
    if ((rc = download_columns(h, current, h->n_surf, group, 1, h->d_current)))
        return rc;
    return download_columns(h, surface_flux, h->n_surf, group, 1, h->d_surfflux);
}

int mocb200_set_sn_xs(mocb200_sweeper *h, int g_begin, int g_count, const double *xs)
{
    int rc = check_groups(h, g_begin, g_count);
    if (rc)
        return rc;
    if (!xs)
        return fail(h, MOCB200_ERR_INVALID, "set_sn_xs: NULL array");
    if (!h->have_corr)
        return fail(h, MOCB200_ERR_STATE, "2D3D correction factors unavailable: %s", h->corr_why.c_str());
    CUDA_TRY(h, cudaSetDevice(h->device));
    if ((rc = upload_columns(h, xs, (int64_t)h->n_plane * h->n_cell_plane, g_begin, g_count, h->d_sn_xs)))
        return rc;
    for (int g = g_begin; g < g_begin + g_count; g++)
        h->have_sn_xs[g] = true;
    return MOCB200_OK;
}

int mocb200_get_corrections(mocb200_sweeper *h, int group, double *alpha, double *beta)
{
    int rc = check_groups(h, group, 1);
    if (rc)
        return rc;
    if (!alpha || !beta)
        return fail(h, MOCB200_ERR_INVALID, "get_corrections: NULL array");
    if (!h->d_alpha || group < h->corr_g_begin || group >= h->corr_g_begin + h->corr_g_count)
        return fail(h, MOCB200_ERR_STATE, "no correction factors of group %d on the device", group);
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const size_t ncp    = (size_t)h->n_cell_plane;
    const size_t n_cell = (size_t)h->n_plane * ncp;
    const size_t grel   = (size_t)(group - h->corr_g_begin);
    // only this handle's macroplanes: one contiguous cell range per angle
    const size_t c0 = (size_t)h->plane_begin * ncp, cn = (size_t)(h->plane_end - h->plane_begin) * ncp;
    for (int ia = 0; ia < 2 * h->n_ang; ia++) {
        const size_t o = (grel * 2 * h->n_ang + ia) * n_cell + c0;
        CUDA_TRY(h, cudaMemcpyAsync(alpha + 2 * ((size_t)ia * n_cell + c0), h->d_alpha + 2 * o, 2 * cn * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaMemcpyAsync(beta + (size_t)ia * n_cell + c0, h->d_beta + o, cn * sizeof(double),
                                    cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return MOCB200_OK;
}

int mocb200_get_stats(const mocb200_sweeper *h, mocb200_stats *out)
{
    if (!h || !out)
        return MOCB200_ERR_INVALID;
    *out = h->stats;
    return MOCB200_OK;
}

int mocb200_last_sweep_ms(mocb200_sweeper *h, double *ms)
{
    if (!h || !ms)
        return MOCB200_ERR_INVALID;
    if (!h->ev_valid)
        return fail(h, MOCB200_ERR_STATE, "no sweep has been timed yet");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaEventSynchronize(h->ev1));
    float f = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&f, h->ev0, h->ev1));
    *ms = f;
    return MOCB200_OK;
}

int mocb200_set_timing(mocb200_sweeper *h, int enabled)
{
    if (!h)
        return MOCB200_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->timing  = enabled != 0;
    h->ev_used = 0;
    return MOCB200_OK;
}

int mocb200_get_timing(mocb200_sweeper *h, double *sweep_ms, int64_t *inner_sweeps)
{
    if (!h || !sweep_ms || !inner_sweeps)
        return MOCB200_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    double tot = 0.0;
    for (size_t i = 0; i < h->ev_used; i++) {
        float f = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&f, h->ev_pool[i].first, h->ev_pool[i].second));
        tot += f;
    }
    int64_t inners = 0; // an event pair of the persistent path brackets several inners
    for (size_t i = 0; i < h->ev_used; i++)
        inners += h->ev_inners[i];
    *sweep_ms     = tot;
    *inner_sweeps = inners;
    h->ev_used    = 0;
    return MOCB200_OK;
}

} // extern "C"
