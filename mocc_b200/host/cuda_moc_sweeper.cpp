#include "cuda_moc_sweeper.hpp"

#include <algorithm>
#include <sstream>

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
    opt.device              = cu.attribute("device").as_int(0);
    opt.max_polar           = cu.attribute("max_polar").as_int(0);
    group_batch_            = cu.attribute("group_batch").as_bool(false);
    std::string kernel      = cu.attribute("kernel").as_string("track");
    if (kernel == "track")
        opt.kernel = MOCB200_KERNEL_TRACK;
    else if (kernel == "item")
        opt.kernel = MOCB200_KERNEL_ITEM;
    else
        throw EXCEPT("Unrecognized <cuda kernel=...> option.");
    if (allow_splitting_ && group_batch_)
        Warn("group_batch with tl_splitting re-uploads the split cross sections every sweep.");

    // ---- flatten the ray data once and hand it to the device ----
    n_macroplane_ = (int)macroplane_unique_ids_.size();
    std::vector<double> vol(vol_.begin(), vol_.end());
    FlatProblem fp = flatten(mesh_, rays_, macroplane_unique_ids_, first_reg_macroplane_, vol.data(), (int)n_reg_,
                             (int)n_group_);
    n_bc_ = fp.bc_per_group;
    if (n_bc_ * (int)n_group_ != boundary_[0].size())
        throw EXCEPT("Flattened boundary layout does not match BoundaryCondition storage.");
    mocb200_problem prob = fp.view();
    int rc               = mocb200_create(&prob, &opt, &dev_);
    if (rc != MOCB200_OK) {
        std::stringstream msg;
        msg << "mocb200_create failed (" << rc << "): " << mocb200_last_error(nullptr);
        throw EXCEPT(msg.str());
    }
    mocb200_get_stats(dev_, &stats_);
    LogFile << "B200 MoC sweeper: " << fp.n_seg_reference << " segments (" << stats_.unique_segments
            << " resident after polar sharing), " << stats_.device_bytes / (1024.0 * 1024.0) << " MiB on device"
            << std::endl;

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
    col_.resize(std::max<size_t>(n_reg_, (size_t)n_bc_));
    cur_.resize(mesh_.n_surf());
    sflux_.resize(mesh_.n_surf());

    timer_init_.toc();
    timer_.toc();
}

CudaMoCSweeper::~CudaMoCSweeper()
{
    if (dev_)
        mocb200_destroy(dev_);
}

void CudaMoCSweeper::check(int rc, const char *what) const
{
    if (rc != MOCB200_OK) {
        std::stringstream msg;
        msg << what << " failed (" << rc << "): " << mocb200_last_error(dev_);
        throw EXCEPT(msg.str());
    }
}

const mocb200_stats &CudaMoCSweeper::device_stats()
{
    mocb200_get_stats(dev_, &stats_);
    return stats_;
}

// Host state of one group -> device: cross sections (when they can have changed), the
// one-group source, the current scalar flux and the incoming boundary flux.
void CudaMoCSweeper::upload_group(int group)
{
    // ExpandedXS::expand as MoCSweeper::sweep does it (moc_sweeper.cpp:197)
    xstr_.expand(group, split_);
    if (!xs_uploaded_[group] || allow_splitting_) {
        std::copy(xstr_.xs().begin(), xstr_.xs().end(), col_.begin());
        check(mocb200_set_xs(dev_, group, 1, col_.data(), &xstr_true_fsr_[(size_t)group * n_reg_],
                             &xs_self_fsr_[(size_t)group * n_reg_]),
              "mocb200_set_xs");
        xs_uploaded_[group] = true;
    }
    const VectorX &src = source_->get();
    check(mocb200_set_source(dev_, group, 1, src.data()), "mocb200_set_source");
    for (int ireg = 0; ireg < (int)n_reg_; ireg++)
        col_[ireg] = flux_(ireg, group);
    check(mocb200_set_flux(dev_, group, 1, col_.data()), "mocb200_set_flux");
    for (int ip = 0; ip < n_macroplane_; ip++)
        check(mocb200_set_boundary(dev_, ip, group, 1, boundary_[ip].get_boundary(group, 0).second),
              "mocb200_set_boundary");
}

// Device results of one group -> host objects the rest of MOCC reads.
void CudaMoCSweeper::download_group(int group, int tally)
{
    check(mocb200_get_flux(dev_, group, 1, col_.data()), "mocb200_get_flux");
    for (int ireg = 0; ireg < (int)n_reg_; ireg++)
        flux_(ireg, group) = col_[ireg];
    for (int ip = 0; ip < n_macroplane_; ip++)
        check(mocb200_get_boundary(dev_, ip, group, 1, boundary_[ip].get_boundary(group, 0).second),
              "mocb200_get_boundary");
    if (tally == MOCB200_TALLY_CURRENT) {
        // moc_sweeper.cpp:208-215: zero the radial data, tally, flag; the raw device tallies
        // then go through the reference's own post_sweep (sub-plane expansion and division
        // by the surface area, moc_current_worker.hpp:272-318) so every quirk is kept.
        coarse_data_->zero_data_radial(group);
        check(mocb200_get_coarse(dev_, group, cur_.data(), sflux_.data()), "mocb200_get_coarse");
        for (int ip = 0; ip < n_macroplane_; ip++) {
            for (int s = mesh_.plane_surf_xy_begin(ip); s < (int)mesh_.plane_surf_end(ip); s++) {
                coarse_data_->current(s, group)      = cur_[s];
                coarse_data_->surface_flux(s, group) = sflux_[s];
            }
        }
        moc::Current cw(coarse_data_, &mesh_);
        cw.set_group(group);
        cw.post_sweep();
        coarse_data_->set_has_radial_data(true);
    }
    post_group(group);
}

void CudaMoCSweeper::sweep(int group)
{
    assert(source_);
    timer_.tic();
    timer_sweep_.tic();

    flux_1g_.reference(flux_(blitz::Range::all(), group));
    upload_group(group);
    const int tally = tally_mode();
    if (!group_batch_) {
        check(mocb200_sweep(dev_, group, 1, (int)n_inner_, tally, 0), "mocb200_sweep");
        download_group(group, tally);
        double ms = 0.0;
        if (mocb200_last_sweep_ms(dev_, &ms) == MOCB200_OK)
            device_sweep_ms_ += ms * n_inner_; // last inner timed; inners are alike
    } else if (group == (int)n_group_ - 1) {
        check(mocb200_sweep(dev_, 0, (int)n_group_, (int)n_inner_, tally, 0), "mocb200_sweep");
        for (int ig = 0; ig < (int)n_group_; ig++)
            download_group(ig, tally);
        double ms = 0.0;
        if (mocb200_last_sweep_ms(dev_, &ms) == MOCB200_OK)
            device_sweep_ms_ += ms * n_inner_;
    }

    timer_.toc();
    timer_sweep_.toc();
}
}
