// cuda_moc_sweeper.hpp -- the drop-in: a TransportSweeper that runs MOCC's 2-D MoC
// transport sweep on a B200 through the C ABI of include/mocc_b200.h.
//
// Selected from the XML input with <sweeper type="moc_cuda" ...> (see
// transport_sweeper_factory.cpp in this directory). It derives from the reference's
// moc::MoCSweeper (src/sweepers/moc/moc_sweeper.hpp:36-234) and overrides exactly the hot
// path, MoCSweeper::sweep (moc_sweeper.cpp:189-225): ray tracing, boundary-condition
// objects, pin homogenisation/prolongation, transverse leakage, fission source and output
// are inherited unchanged, so the rest of MOCC (EigenSolver, FixedSourceSolver, CMFD, 2D3D)
// runs as is. Host-visible state (flux_, boundary_, CoarseData) is refreshed after every
// sweep(group) because the callers read it between sweeps (transport_sweeper.hpp:139-150,
// source.cpp:26-32).
//
// Options live on a <cuda> child of <sweeper> so that the reference's attribute validation
// (moc_sweeper.cpp:54-57, util/validate_input.cpp:24-42) stays quiet:
//   <cuda device="0" group_batch="f" kernel="auto" max_polar="2"/>      one GPU
//   <cuda devices="0,1,2,3" .../>                                       macroplanes sharded over several GPUs
//     group_batch="t": sweep(0..ng-2) only stage their sources; sweep(ng-1) sweeps all groups
//                      in one batch (Jacobi instead of Gauss-Seidel in energy: same converged
//                      answer, different iteration path).
#pragma once

#include <string>
#include <vector>

#include "pugixml.hpp"

#include "core/core_mesh.hpp"
#include "sweepers/moc/moc_sweeper.hpp"

#include "sweepers/cmdo/correction_data.hpp"

#include "mocc_b200.h"
#include "xs_updater.hpp"

namespace mocc_b200 {

// The Source CudaMoCSweeper::create_source hands to MOCC's FixedSourceSolver. The solver still
// calls initialize_group / fission / in_scatter before every sweep(group) (fixed_source_solver.cpp:102-117), but the
// O(n_reg G) host loops of Source::fission and Source::in_scatter (source.cpp:64-112) are not run: the calls are
// recorded and CudaMoCSweeper::sweep has the device build the same source from the flux resident there
// (mocb200_set_fission_source + mocb200_build_source, bit-identical). With an external source, or when anything adds
// to the host source behind the device's back, the reference's host path is used as it is.
class DeviceSource : public mocc::SourceIsotropic {
public:
    DeviceSource(int nreg, const mocc::XSMesh *xs_mesh, const mocc::ArrayB2 &flux, bool check)
        : mocc::SourceIsotropic(nreg, xs_mesh, flux), check_(check)
    {
    }
    void initialize_group(int ig) override;
    void fission(const mocc::ArrayB1 &fs, int ig) override;
    void in_scatter(size_t ig) override;
    // true: the source of `group` is to be built on the device (fission source: fs(), nullptr = none)
    bool deferred(int group) const
    {
        return !has_external_ && group_ == group && scattered_;
    }
    const mocc::ArrayB1 *fs() const
    {
        return fs_;
    }
    // false: Source::get() does not hold this group's fission / in-scatter source (they were left to the device)
    bool host_built() const
    {
        return has_external_ || check_;
    }
    bool check() const // MOCB200_CHECK_DEVICE_SOURCES: the host builds its source too, the sweeper compares every bit
    {
        return check_;
    }

private:
    const mocc::ArrayB1 *fs_ = nullptr;
    int group_               = -1;
    bool scattered_          = false;
    bool check_              = false;
};

class CudaMoCSweeper : public mocc::moc::MoCSweeper {
public:
    CudaMoCSweeper(const pugi::xml_node &input, const mocc::CoreMesh &mesh);
    ~CudaMoCSweeper();

    void sweep(int group) override;

    // moc_sweeper.hpp:90-95 (`final` there: nofinal_moc_sweeper.hpp): a DeviceSource when sources are built on the
    // device (<cuda device_sources="t">, the default of type="moc_cuda" without group batching)
    mocc::UP_Source_t create_source(const pugi::xml_node &input) const override;
    // the host rewrites the flux (moc_sweeper.cpp:234-253, 360-433): the device copy is refreshed before the next source
    void initialize() override;
    mocc::real_t set_pin_flux_1g(int group, const mocc::ArrayB1 &pin_flux,
                                 mocc::MeshTreatment treatment = mocc::MeshTreatment::PIN_PLANE) override;

    // Device time spent in transport-sweep kernels / number of C-ABI sweeps so far
    double device_sweep_ms() const
    {
        return device_sweep_ms_;
    }
    const mocb200_stats &device_stats();

protected:
    // What the last inner iteration of a sweep tallies; the 2D3D variant overrides this
    virtual int tally_mode() const
    {
        return coarse_data_ ? MOCB200_TALLY_CURRENT : MOCB200_TALLY_NONE;
    }
    // Hook between the first n_inner-1 inners and the last one (2D3D refreshes the Sn cross sections there)
    virtual bool split_last_inner() const
    {
        return false;
    }
    virtual void before_last_inner(int group)
    {
        (void)group;
    }
    // Hook called after the device results of `group` are back on the host
    virtual void post_group(int group, int tally)
    {
        (void)group;
        (void)tally;
    }
    void download_flux(int group);

    // One C-ABI handle per GPU, each owning a contiguous range of macroplanes (planes are independent
    // inside a sweep, moc_sweeper_kernel.inc.hpp:51-153); a single device owns all of them.
    struct Part {
        mocb200_sweeper *h = nullptr;
        int device = 0, plane_begin = 0, plane_end = 0, reg_lo = 0, reg_hi = 0;
    };
    void check(const Part &p, int rc, const char *what) const;
    template <class Fn> void for_each_part(Fn &&fn); // fn(part, index), one host thread per part
    void upload_group(int group);
    void download_group(int group, int tally);

    std::vector<Part> parts_;
    bool group_batch_ = false;
    // device-side sources
    bool device_sources_ = false; // <cuda device_sources="t|f"> (default t)
    virtual bool device_sources_allowed() const // the 2D3D variant adds transverse leakage to the host source
    {
        return true;
    }
    bool source_xs_sent_ = false, fs_sent_ = false;
    long n_source_checks_ = 0;
    std::vector<bool> flux_stale_; // per group: the host has rewritten the column since the device last had it
    std::vector<double> flux_all_;
    void upload_source_tables();
    int n_bc_             = 0; // boundary values per group per plane
    int n_macroplane_     = 0;
    // per-FSR cross sections, [n_group][n_reg]
    std::vector<double> xstr_true_fsr_; // un-split transport XS (source normalisation)
    std::vector<double> xs_self_fsr_;   // within-group scattering
    std::vector<bool> xs_uploaded_;
    std::vector<int> plane_xs_offset_; // CurrentCorrections::mplane_offset_
    std::vector<double> col_, cur_, sflux_;
    double device_sweep_ms_ = 0.0;
    std::vector<double> part_ms_; // per device: sum over all sweep(group) calls of the event time of every inner
    // wall-clock split of sweep(): host->device, host work between the inners (2D3D), waiting for the
    // device + device->host, host post-processing; reported at destruction
    double t_upload_ = 0.0, t_host_mid_ = 0.0, t_download_ = 0.0, t_enqueue_ = 0.0;
    long n_sweep_calls_ = 0;
    mocb200_stats stats_{};
};

// The per-plane MoC sweeper of the 2D3D method on the B200: same interface as
// cmdo::MoCSweeper_2D3D (src/sweepers/cmdo/moc_sweeper_2d3d.hpp:25-88), so that the reference's
// PlaneSweeper_2D3D can hold it in place of the CPU class (plane_sweeper_2d3d_cuda.cpp).
// The last inner iteration of every sweep(group) tallies coarse currents AND the CDD correction
// factors alpha/beta (cmdo::CurrentCorrections, correction_worker.hpp:36-290) on the device.
class CudaMoCSweeper2D3D : public CudaMoCSweeper {
public:
    CudaMoCSweeper2D3D(const pugi::xml_node &input, const mocc::CoreMesh &mesh);

    void sweep(int group) override;

    void set_coupling(std::shared_ptr<mocc::CorrectionData> data, mocc::SP_XSMeshHomogenized_t xsmesh,
                      mocc::ExpandedXS &xstr);
    void set_self_coupling();
    void output(mocc::H5Node &node) const override;

protected:
    int tally_mode() const override
    {
        return MOCB200_TALLY_CORRECTIONS;
    }
    bool split_last_inner() const override
    {
        return true;
    }
    bool device_sources_allowed() const override
    {
        return false;
    }
    void before_last_inner(int group) override;
    void post_group(int group, int tally) override;

private:
    std::shared_ptr<mocc::CorrectionData> corrections_;
    std::shared_ptr<mocc::XSMeshHomogenized> sn_xs_mesh_;
    XsUpdater xs_updater_;
    mocc::ExpandedXS xstr_sn_;
    bool internal_coupling_ = false;
    std::vector<double> sn_col_, alpha_, beta_;
    std::vector<std::vector<std::array<mocc::real_t, 3>>> correction_residuals_;
};
}
