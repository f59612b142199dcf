#include "cuda_moc_sweeper.hpp"

#include <cctype>
#include <chrono>
#include <cstring>
#include <cstdio>

#include <algorithm>
#include <array>
#include <cmath>
#include <exception>
#include <sstream>
#include <thread>

#include "core/coarse_data.hpp"
#include "core/source.hpp"
#include "sweepers/moc/moc_current_worker.hpp"
#include "util/error.hpp"
#include "util/files.hpp"

#include "flatten.hpp"

using namespace mocc;

namespace mocc_b200 {

CudaMoCSweeper::CudaMoCSweeper(const pugi::xml_node &input, const CoreMesh &mesh) : moc::MoCSweeper(input, mesh)
{
    LogFile << "Constructing the B200 (CUDA) MoC sweeper" << std::endl;
    timer_.tic();
    timer_init_.tic();

    // ---- options ----
    mocb200_options opt{};
    opt.boundary_update = gauss_seidel_boundary_ ? MOCB200_BOUNDARY_GS : MOCB200_BOUNDARY_JACOBI;
    const pugi::xml_node cu = input.child("cuda");
    std::vector<int> devices;
    {
        std::stringstream ds(cu.attribute("devices").as_string(""));
        std::string tok;
        while (std::getline(ds, tok, ','))
            if (!tok.empty())
                devices.push_back(std::stoi(tok));
        if (devices.empty())
            devices.push_back(cu.attribute("device").as_int(0));
    }
    opt.max_polar           = cu.attribute("max_polar").as_int(0);
    opt.cache_groups        = cu.attribute("cache_groups").as_int(0);
    group_batch_            = cu.attribute("group_batch").as_bool(false);
    device_sources_         = cu.attribute("device_sources").as_bool(true);
    std::string kernel      = cu.attribute("kernel").as_string("auto");
    if (kernel == "auto")
        opt.kernel = MOCB200_KERNEL_AUTO;
    else if (kernel == "track")
        opt.kernel = MOCB200_KERNEL_TRACK;
    else if (kernel == "cached")
        opt.kernel = MOCB200_KERNEL_CACHED;
    else if (kernel == "item")
        opt.kernel = MOCB200_KERNEL_ITEM;
    else if (kernel == "chunk")
        opt.kernel = MOCB200_KERNEL_CHUNK;
    else if (kernel == "rchunk")
        opt.kernel = MOCB200_KERNEL_RCHUNK;
    else
        throw EXCEPT("Unrecognized <cuda kernel=...> option.");
    if (allow_splitting_ && group_batch_)
        Warn("group_batch with tl_splitting re-uploads the split cross sections every sweep.");

    // ---- flatten the ray data once and hand it to the device ----
    n_macroplane_ = (int)macroplane_unique_ids_.size();
    std::vector<double> vol(vol_.begin(), vol_.end());
    FlatProblem fp = flatten(mesh_, rays_, macroplane_unique_ids_, first_reg_macroplane_, vol.data(), (int)n_reg_,
                             (int)n_group_);
    n_bc_            = fp.bc_per_group;
    plane_xs_offset_ = fp.plane_xs_offset;
    if (n_bc_ * (int)n_group_ != boundary_[0].size())
        throw EXCEPT("Flattened boundary layout does not match BoundaryCondition storage.");
    mocb200_problem prob = fp.view();
    // contiguous macroplane ranges, balanced by the segments each plane's ray set holds
    std::vector<double> plane_w(n_macroplane_, 0.0);
    for (int ip = 0; ip < n_macroplane_; ip++) {
        const int u = macroplane_unique_ids_[ip];
        for (int gi = 0; gi < fp.n_geom; gi++) {
            const int64_t t0 = fp.geom_trk_begin[(size_t)u * fp.n_geom + gi], t1 = fp.geom_trk_begin[(size_t)u * fp.n_geom + gi + 1];
            plane_w[ip] += (double)(fp.trk_seg_begin[t1] - fp.trk_seg_begin[t0]);
        }
    }
    const int n_part = std::min<int>((int)devices.size(), n_macroplane_);
    double w_total = 0.0;
    for (double w : plane_w)
        w_total += w;
    int ip = 0;
    double w_done = 0.0;
    for (int k = 0; k < n_part; k++) {
        Part part;
        part.device      = devices[k];
        part.plane_begin = ip;
        const double target = w_total * (k + 1) / n_part;
        while (ip < n_macroplane_ && (ip == part.plane_begin || w_done + 0.5 * plane_w[ip] < target) &&
               n_macroplane_ - ip > n_part - 1 - k) {
            w_done += plane_w[ip];
            ip++;
        }
        if (k == n_part - 1)
            ip = n_macroplane_;
        part.plane_end = ip;
        part.reg_lo    = first_reg_macroplane_[part.plane_begin];
        part.reg_hi    = part.plane_end < n_macroplane_ ? first_reg_macroplane_[part.plane_end] : (int)n_reg_;
        opt.device      = part.device;
        opt.plane_begin = part.plane_begin;
        opt.plane_end   = part.plane_end;
        int rc          = mocb200_create(&prob, &opt, &part.h);
        if (rc != MOCB200_OK) {
            std::stringstream msg;
            msg << "mocb200_create failed (" << rc << "): " << mocb200_last_error(nullptr);
            throw EXCEPT(msg.str());
        }
        parts_.push_back(part);
        part_ms_.push_back(0.0);
        mocb200_set_timing(part.h, 1); // CUDA events around the sweep kernels of every inner (device_sweep_ms)
        mocb200_get_stats(part.h, &stats_);
        LogFile << "B200 MoC sweeper: device " << part.device << " owns macroplanes [" << part.plane_begin << ", "
                << part.plane_end << "): " << stats_.segments_per_sweep << " segments per sweep, "
                << stats_.device_bytes / (1024.0 * 1024.0) << " MiB on device" << std::endl;
    }
    LogFile << "B200 MoC sweeper: " << fp.n_seg_reference << " segments (" << fp.seg_len.size()
            << " resident after polar sharing) on " << parts_.size() << " GPU(s)" << std::endl;

    // ---- per-FSR cross sections the device-side self-scatter source needs ----
    xstr_true_fsr_.assign((size_t)n_group_ * n_reg_, 0.0);
    xs_self_fsr_.assign((size_t)n_group_ * n_reg_, 0.0);
    for (const auto &xsr : *xs_mesh_) {
        for (int ig = 0; ig < (int)n_group_; ig++) {
            const real_t tr = xsr.xsmactr(ig);
            const real_t sc = xsr.xsmacsc().to(ig)[ig];
            for (const int ireg : xsr.reg()) {
                xstr_true_fsr_[(size_t)ig * n_reg_ + ireg] = tr;
                xs_self_fsr_[(size_t)ig * n_reg_ + ireg]   = sc;
            }
        }
    }
    xs_uploaded_.assign(n_group_, false);
    flux_stale_.assign(n_group_, true);
    col_.resize(std::max<size_t>(n_reg_, (size_t)n_bc_));
    cur_.resize(mesh_.n_surf());
    sflux_.resize(mesh_.n_surf());

    timer_init_.toc();
    timer_.toc();
}

CudaMoCSweeper::~CudaMoCSweeper()
{
    if (n_sweep_calls_ > 0) {
        std::printf("CudaMoCSweeper: %ld sweep(group) calls; wall-clock inside them: upload %.3f s, enqueue %.3f s, "
                    "host work between inners %.3f s, device wait + download + post-processing %.3f s; device sweep "
                    "kernels %.3f s\n",
                    n_sweep_calls_, t_upload_, t_enqueue_, t_host_mid_, t_download_, device_sweep_ms_ * 1e-3);
    }
    for (auto &p : parts_)
        if (p.h)
            mocb200_destroy(p.h);
}

// fn(part) for every device part; with several parts one host thread each, so that the blocking
// device->host copies (and the packing into pinned staging) of the devices overlap. The C ABI is
// thread-compatible per handle: every call sets its own device.
template <class Fn> void CudaMoCSweeper::for_each_part(Fn &&fn)
{
    if (parts_.size() == 1) {
        fn(parts_[0], 0);
        return;
    }
    std::vector<std::exception_ptr> err(parts_.size());
    std::vector<std::thread> th;
    for (size_t i = 0; i < parts_.size(); i++)
        th.emplace_back([&, i]() {
            try {
                fn(parts_[i], i);
            } catch (...) {
                err[i] = std::current_exception();
            }
        });
    for (auto &t : th)
        t.join();
    for (auto &e : err)
        if (e)
            std::rethrow_exception(e);
}

void CudaMoCSweeper::check(const Part &p, int rc, const char *what) const
{
    if (rc != MOCB200_OK) {
        std::stringstream msg;
        msg << what << " failed on device " << p.device << " (" << rc << "): " << mocb200_last_error(p.h);
        throw EXCEPT(msg.str());
    }
}

const mocb200_stats &CudaMoCSweeper::device_stats()
{
    mocb200_get_stats(parts_[0].h, &stats_);
    return stats_;
}

// ---- device-side sources ----
void DeviceSource::initialize_group(int ig)
{
    mocc::Source::initialize_group(ig); // external source or zero (source.cpp:41-54)
    group_     = ig;
    fs_        = nullptr;
    scattered_ = false;
}

void DeviceSource::fission(const mocc::ArrayB1 &fs, int ig)
{
    fs_ = &fs; // chi_g * fs is added on the device (source.cpp:64-80)
    if (has_external_ || check_)
        mocc::Source::fission(fs, ig);
}

void DeviceSource::in_scatter(size_t ig)
{
    scattered_ = true; // sum over g' != g of Sigma_s(g' -> g) flux_g' is added on the device (source.cpp:86-112)
    if (has_external_ || check_)
        mocc::Source::in_scatter(ig);
}

mocc::UP_Source_t CudaMoCSweeper::create_source(const pugi::xml_node &input) const
{
    if (!device_sources_ || group_batch_ || !device_sources_allowed())
        return mocc::moc::MoCSweeper::create_source(input);
    // SourceFactory (source_factory.cpp:26-69) with the P0 source replaced
    std::string scat = input.attribute("scattering").value();
    for (auto &c : scat)
        c = (char)std::tolower(c);
    if (input.empty() || scat != "p0")
        return mocc::moc::MoCSweeper::create_source(input); // the reference's own error handling
    mocc::UP_Source_t source(
        new DeviceSource((int)n_reg_, xs_mesh_.get(), this->flux(), std::getenv("MOCB200_CHECK_DEVICE_SOURCES") != nullptr));
    source->add_external(input);
    return source;
}

void CudaMoCSweeper::initialize()
{
    mocc::moc::MoCSweeper::initialize();
    flux_stale_.assign(n_group_, true);
}

mocc::real_t CudaMoCSweeper::set_pin_flux_1g(int group, const mocc::ArrayB1 &pin_flux, mocc::MeshTreatment treatment)
{
    flux_stale_[group] = true;
    return mocc::moc::MoCSweeper::set_pin_flux_1g(group, pin_flux, treatment);
}

// cross sections by cross-section-mesh region for the device's fission / in-scatter sources
void CudaMoCSweeper::upload_source_tables()
{
    const int ng = (int)n_group_;
    std::vector<int32_t> fsr_mat(n_reg_, 0), band;
    std::vector<double> nf, ch, scat;
    int m = 0;
    for (const auto &xsr : *xs_mesh_) {
        for (int g = 0; g < ng; g++) {
            nf.push_back(xsr.xsmacnf(g));
            ch.push_back(xsr.xsmacch(g));
            const mocc::ScatteringRow &row = xsr.xsmacsc().to(g);
            band.push_back(row.min_g);
            band.push_back(row.max_g);
            for (int gf = 0; gf < ng; gf++)
                scat.push_back(gf >= row.min_g && gf <= row.max_g ? row[gf] : 0.0);
        }
        for (const int ireg : xsr.reg())
            fsr_mat[ireg] = m;
        m++;
    }
    for_each_part([&](const Part &p, size_t) {
        check(p, mocb200_set_source_xs(p.h, m, fsr_mat.data(), nf.data(), ch.data(), scat.data(), band.data()),
              "mocb200_set_source_xs");
    });
    source_xs_sent_ = true;
}

// Host state of one group -> device: cross sections (when they can have changed), the
// one-group source, the current scalar flux and the incoming boundary flux.
void CudaMoCSweeper::upload_group(int group)
{
    // ExpandedXS::expand as MoCSweeper::sweep does it (moc_sweeper.cpp:197)
    xstr_.expand(group, split_);
    const bool send_xs = !xs_uploaded_[group] || allow_splitting_;
    if (send_xs)
        std::copy(xstr_.xs().begin(), xstr_.xs().end(), col_.begin());
    const VectorX &src = source_->get();
    if (send_xs)
        for_each_part([&](const Part &p, size_t) {
            check(p, mocb200_set_xs(p.h, group, 1, col_.data(), &xstr_true_fsr_[(size_t)group * n_reg_],
                                    &xs_self_fsr_[(size_t)group * n_reg_]),
                  "mocb200_set_xs");
        });
    xs_uploaded_[group] = true;
    std::vector<const double *> bc(n_macroplane_, nullptr);
    for (int ip = 0; ip < n_macroplane_; ip++)
        bc[ip] = boundary_[ip].get_boundary(group, 0).second;
    DeviceSource *ds = dynamic_cast<DeviceSource *>(source_);
    if (ds && ds->deferred(group)) {
        // The device builds this group's source from the flux resident there: the host sends the fission source
        // (once per outer: the solver recomputes it before group 0, eigen_solver.cpp:229-231) and the flux columns it
        // has rewritten since the device last had them (after initialize() and after the CMFD prolongation: all of
        // them, once per outer; none inside a fixed-source iteration).
        if (!source_xs_sent_)
            upload_source_tables();
        const int ng = (int)n_group_;
        for (int g = 0; g < ng; g++) {
            if (!flux_stale_[g])
                continue;
            for (int ireg = 0; ireg < (int)n_reg_; ireg++)
                col_[ireg] = flux_(ireg, g);
            for_each_part([&](const Part &p, size_t) { check(p, mocb200_set_flux(p.h, g, 1, col_.data()), "mocb200_set_flux"); });
            flux_stale_[g] = false;
        }
        if (group == 0 || !fs_sent_) {
            flux_all_.assign(n_reg_, 0.0); // no fission source (fixed-source problem): zero
            if (ds->fs())
                std::copy(ds->fs()->begin(), ds->fs()->end(), flux_all_.begin());
            for_each_part([&](const Part &p, size_t) {
                check(p, mocb200_set_fission_source(p.h, flux_all_.data()), "mocb200_set_fission_source");
            });
            fs_sent_ = true;
        }
        for_each_part([&](const Part &p, size_t) {
            check(p, mocb200_build_source(p.h, group, 1), "mocb200_build_source");
            check(p, mocb200_set_sweep_inputs(p.h, group, nullptr, nullptr, bc.data()), "mocb200_set_sweep_inputs");
        });
        if (ds->check()) { // the host built its source too: every bit must agree
            std::vector<double> dev(n_reg_);
            const Part &p = parts_[0];
            check(p, mocb200_get_source(p.h, group, 1, dev.data()), "mocb200_get_source");
            for (int ireg = p.reg_lo; ireg < p.reg_hi; ireg++)
                if (std::memcmp(&dev[ireg], &src.data()[ireg], sizeof(double)) != 0) {
                    std::stringstream msg;
                    msg << "device-built source of group " << group << " differs from the host's in FSR " << ireg << ": "
                        << std::setprecision(17) << dev[ireg] << " vs " << src.data()[ireg];
                    throw EXCEPT(msg.str());
                }
            n_source_checks_++;
        }
        return;
    }
    if (ds && !ds->host_built()) {
        // sweep(group) without initialize_group / in_scatter for THIS group right before it (FixedSourceSolver::step
        // always does both): the host array holds no fission / in-scatter source to fall back to
        std::stringstream msg;
        msg << "CudaMoCSweeper::sweep(" << group << "): the source of this group was not prepared (Source::initialize_group "
               "and in_scatter must precede the sweep when sources are built on the device; <cuda device_sources=\"f\"/> "
               "restores host-built sources)";
        throw EXCEPT(msg.str());
    }
    for (int ireg = 0; ireg < (int)n_reg_; ireg++)
        col_[ireg] = flux_(ireg, group);
    flux_stale_[group] = false;
    // source, flux and incoming boundary flux in one staged copy per device (no synchronisation)
    for_each_part([&](const Part &p, size_t) {
        check(p, mocb200_set_sweep_inputs(p.h, group, src.data(), col_.data(), bc.data()), "mocb200_set_sweep_inputs");
    });
}

void CudaMoCSweeper::download_flux(int group)
{
    // every handle writes the FSR range of its own macroplanes (the rest of col_ is left alone)
    for_each_part([&](const Part &p, size_t) {
        check(p, mocb200_get_sweep_results(p.h, group, col_.data(), nullptr, nullptr, nullptr), "mocb200_get_sweep_results");
        for (int ireg = p.reg_lo; ireg < p.reg_hi; ireg++)
            flux_(ireg, group) = col_[ireg];
    });
}

// Device results of one group -> host objects the rest of MOCC reads.
void CudaMoCSweeper::download_group(int group, int tally)
{
    // flux, outgoing boundary flux and the raw coarse tallies in one staged copy and ONE synchronisation per device
    std::vector<double *> bc(n_macroplane_, nullptr);
    for (int ip = 0; ip < n_macroplane_; ip++)
        bc[ip] = const_cast<double *>(boundary_[ip].get_boundary(group, 0).second);
    const bool coarse = tally != MOCB200_TALLY_NONE;
    if (coarse)
        coarse_data_->zero_data_radial(group); // moc_sweeper.cpp:208-215: zero the radial data, tally, flag
    // the devices are drained concurrently (one host thread per handle): flux and boundary flux land in
    // disjoint ranges of the host arrays, the coarse tallies in per-device buffers
    const size_t n_surf = mesh_.n_surf();
    if (coarse && cur_.size() < parts_.size() * n_surf) {
        cur_.resize(parts_.size() * n_surf);
        sflux_.resize(parts_.size() * n_surf);
    }
    for_each_part([&](const Part &p, size_t i) {
        double *cur = cur_.data() + i * n_surf, *sfl = sflux_.data() + i * n_surf;
        check(p, mocb200_get_sweep_results(p.h, group, col_.data(), bc.data(), coarse ? cur : nullptr, coarse ? sfl : nullptr),
              "mocb200_get_sweep_results");
        for (int ireg = p.reg_lo; ireg < p.reg_hi; ireg++)
            flux_(ireg, group) = col_[ireg];
        if (coarse) {
            // the raw device tallies then go through the reference's own post_sweep (sub-plane expansion and
            // division by the surface area, moc_current_worker.hpp:272-318) so every quirk is kept
            for (int ip = p.plane_begin; ip < p.plane_end; ip++) {
                for (int s = mesh_.plane_surf_xy_begin(ip); s < (int)mesh_.plane_surf_end(ip); s++) {
                    coarse_data_->current(s, group)      = cur[s];
                    coarse_data_->surface_flux(s, group) = sfl[s];
                }
            }
        }
        // device time of this sweep(group): CUDA events around the sweep kernels of every inner (the handles
        // run concurrently: the slowest one counts)
        double ms = 0.0;
        int64_t n  = 0;
        if (mocb200_get_timing(p.h, &ms, &n) == MOCB200_OK)
            part_ms_[i] += ms;
    });
    if (coarse) {
        moc::Current cw(coarse_data_, &mesh_);
        cw.set_group(group);
        cw.post_sweep();
        coarse_data_->set_has_radial_data(true);
    }
    post_group(group, tally);
}

void CudaMoCSweeper::sweep(int group)
{
    assert(source_);
    timer_.tic();
    timer_sweep_.tic();

    using clk = std::chrono::steady_clock;
    auto since = [](clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); };
    n_sweep_calls_++;
    flux_1g_.reference(flux_(blitz::Range::all(), group));
    auto t0 = clk::now();
    upload_group(group);
    t_upload_ += since(t0);
    const int tally = tally_mode();
    auto run = [&](int g0, int gc) {
        // every mocb200_sweep only enqueues work on its device's stream: the GPUs sweep concurrently
        if (split_last_inner() && tally != MOCB200_TALLY_NONE) {
            // the host refreshes data between the plain inners and the tallying one
            auto t1 = clk::now();
            if (n_inner_ > 1) {
                for (const Part &p : parts_)
                    check(p, mocb200_sweep(p.h, g0, gc, (int)n_inner_ - 1, MOCB200_TALLY_NONE, 0), "mocb200_sweep");
                t_enqueue_ += since(t1);
                t1 = clk::now();
                for (int ig = g0; ig < g0 + gc; ig++)
                    download_flux(ig);
                t_download_ += since(t1);
            }
            t1 = clk::now();
            for (int ig = g0; ig < g0 + gc; ig++)
                before_last_inner(ig);
            t_host_mid_ += since(t1);
            t1 = clk::now();
            for (const Part &p : parts_)
                check(p, mocb200_sweep(p.h, g0, gc, 1, tally, 0), "mocb200_sweep");
            t_enqueue_ += since(t1);
        } else {
            auto t1 = clk::now();
            for (const Part &p : parts_)
                check(p, mocb200_sweep(p.h, g0, gc, (int)n_inner_, tally, 0), "mocb200_sweep");
            t_enqueue_ += since(t1);
        }
        auto t2 = clk::now();
        for (int ig = g0; ig < g0 + gc; ig++)
            download_group(ig, tally);
        t_download_ += since(t2);
        device_sweep_ms_ = *std::max_element(part_ms_.begin(), part_ms_.end()); // slowest GPU
    };
    if (!group_batch_)
        run(group, 1);
    else if (group == (int)n_group_ - 1)
        run(0, (int)n_group_);

    timer_.toc();
    timer_sweep_.toc();
}

// ---------------------------------------------------------------------------------------------
CudaMoCSweeper2D3D::CudaMoCSweeper2D3D(const pugi::xml_node &input, const CoreMesh &mesh)
    : CudaMoCSweeper(input, mesh), correction_residuals_(n_group_)
{
    LogFile << "Constructing the B200 (CUDA) 2D3D MoC sweeper" << std::endl;
    const size_t n_cell = (size_t)n_macroplane_ * mesh_.nx() * mesh_.ny();
    sn_col_.resize(n_cell);
    alpha_.resize(n_cell * ang_quad_.ndir() / 2 * 2);
    beta_.resize(n_cell * ang_quad_.ndir() / 2);
}

void CudaMoCSweeper2D3D::set_coupling(std::shared_ptr<CorrectionData> data, SP_XSMeshHomogenized_t xsmesh,
                                      ExpandedXS &xstr)
{
    if (corrections_ || sn_xs_mesh_)
        throw EXCEPT("Correction data already assigned.");
    corrections_ = data;
    sn_xs_mesh_  = xsmesh;
    xstr_sn_     = xstr;
}

void CudaMoCSweeper2D3D::set_self_coupling()
{
    internal_coupling_ = true;
    corrections_ =
        std::shared_ptr<CorrectionData>(new CorrectionData(mesh_, ang_quad_.ndir() / 2, xs_mesh_->n_group()));
    sn_xs_mesh_ = this->get_homogenized_xsmesh();
    sn_xs_mesh_->set_flux(flux_);
    xstr_sn_ = ExpandedXS(sn_xs_mesh_.get());
}

void CudaMoCSweeper2D3D::sweep(int group)
{
    assert(sn_xs_mesh_);
    if (!coarse_data_)
        throw EXCEPT("2D3D MoC sweeper needs coarse data to collect "
                     "calculate correction factors. Try enabling CMFD.");
    n_sweep_++;
    n_sweep_inner_ += n_inner_;
    CudaMoCSweeper::sweep(group);
}

// moc_sweeper_2d3d.cpp:85-88: right before the tallying inner the Sn mesh is re-homogenised with the
// current scalar flux; beta divides by that homogenised XS (correction_worker.cpp:84-93).
void CudaMoCSweeper2D3D::before_last_inner(int group)
{
    if (group == 0 || !group_batch_)
        xs_updater_.update(*sn_xs_mesh_); // = sn_xs_mesh_->update(), pins spread over the host threads
    xstr_sn_.expand(group);
    const int ncp = mesh_.nx() * mesh_.ny();
    for (int ip = 0; ip < n_macroplane_; ip++)
        for (int ic = 0; ic < ncp; ic++)
            sn_col_[(size_t)ip * ncp + ic] = xstr_sn_[ic + plane_xs_offset_[ip]];
    for_each_part([&](const Part &p, size_t) { check(p, mocb200_set_sn_xs(p.h, group, 1, sn_col_.data()), "mocb200_set_sn_xs"); });
}

// Device correction factors -> CorrectionData, with the residual bookkeeping of
// calculate_corrections (correction_worker.cpp:112-151) in the reference's plane/angle/cell order.
void CudaMoCSweeper2D3D::post_group(int group, int tally)
{
    if (tally != MOCB200_TALLY_CORRECTIONS)
        return;
    for_each_part([&](const Part &p, size_t) { // each handle fills the cells of its own macroplanes
        check(p, mocb200_get_corrections(p.h, group, alpha_.data(), beta_.data()), "mocb200_get_corrections");
    });
    const int ncp       = mesh_.nx() * mesh_.ny();
    const size_t n_cell = (size_t)n_macroplane_ * ncp;
    const int n_ang     = ang_quad_.ndir() / 4; // sweep angles (octants 1-2)
    // one pass over (plane, angle, cell): 4.7 M entries per group on C5G7-3D, spread over the host threads (with one
    // thread the residual sums run in the reference's plane / angle / cell order)
    real_t r0 = 0.0, r1 = 0.0, r2 = 0.0;
#pragma omp parallel for collapse(2) reduction(+ : r0, r1, r2) schedule(static)
    for (int ip = 0; ip < n_macroplane_; ip++) {
        for (int a = 0; a < n_ang; a++) {
            const int cell_offset = mesh_.coarse_cell_offset(ip);
            for (int ic = 0; ic < ncp; ic++) {
                const int icc = ic + cell_offset;
                for (int iang : {a, (int)ang_quad_.reverse(a)}) {
                    const real_t ax = alpha_[((size_t)iang * n_cell + icc) * 2 + 0];
                    const real_t ay = alpha_[((size_t)iang * n_cell + icc) * 2 + 1];
                    const real_t b  = beta_[(size_t)iang * n_cell + icc];
                    real_t e        = ax - corrections_->alpha(icc, iang, group, Normal::X_NORM);
                    r0 += e * e;
                    e = ay - corrections_->alpha(icc, iang, group, Normal::Y_NORM);
                    r1 += e * e;
                    e = b - corrections_->beta(icc, iang, group);
                    r2 += e * e;
                    corrections_->alpha(icc, iang, group, Normal::X_NORM) = ax;
                    corrections_->alpha(icc, iang, group, Normal::Y_NORM) = ay;
                    corrections_->beta(icc, iang, group)                  = b;
                }
            }
        }
    }
    std::array<real_t, 3> resid = {{r0, r1, r2}};
    correction_residuals_[group].push_back({{std::sqrt(resid[0]), std::sqrt(resid[1]), std::sqrt(resid[2])}});
}

void CudaMoCSweeper2D3D::output(H5Node &node) const
{
    LogFile << "MoC Sweeper 2D3D (B200) output:" << std::endl;
    LogFile << "    Number of sweeps, outer: " << n_sweep_ << std::endl;
    LogFile << "    Number of sweeps, inner: " << n_sweep_inner_ << std::endl;
    MoCSweeper::output(node);
    if (internal_coupling_) {
        corrections_->output(node);
        sn_xs_mesh_->update();
        sn_xs_mesh_->output(node);
    }
    auto residual_group = node.create_group("correction_residual");
    for (int ig = 0; ig < (int)n_group_; ig++) {
        auto g = residual_group.create_group(std::to_string(ig + 1));
        VecF ax, ay, b;
        for (const auto &d : correction_residuals_[ig]) {
            ax.push_back(d[0]), ay.push_back(d[1]), b.push_back(d[2]);
        }
        g.write("alpha_x", ax);
        g.write("alpha_y", ay);
        g.write("beta", b);
    }
}
}
