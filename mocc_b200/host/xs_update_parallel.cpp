// xs_update_parallel.cpp -- XSMeshHomogenized::update() for the 2D3D plugin, restated on flat tables.
//
// MoCSweeper_2D3D::sweep re-homogenises the Sn cross-section mesh before the last inner of EVERY group
// (moc_sweeper_2d3d.cpp:85-88 -> XSMeshHomogenized::update, xs_mesh_homogenized.cpp:176-197 ->
// homogenize_region_flux, :261-362). The reference does that serially and copies the material library for every
// pin and a Material for every cross-section region and group (:281, :293, :316); on C5G7-sized planes that is
// what a sweep(group) call costs once the MoC sweep itself runs on the GPU (profiles/r1: 0.48 s of 0.54 s per
// call on the 9-plane C5G7 3-D case).
//
// This file restates the flux-volume weighting of homogenize_region_flux on tables flattened once -- per pin the
// areas and material of its regions, per material the cross sections and scattering rows -- with the SAME order
// of floating-point operations, pins spread over the host threads (every pin writes only its own
// XSMeshRegion). Results are bit-identical to the reference routine; MOCB200_CHECK_XS_UPDATE=1 runs the
// reference's routine next to it and throws on the first differing bit (tests/test_gpu_plugin.py).
//
// The class keeps its regions, flux pointer and state counter non-public; this translation unit alone sees the
// reference headers with that access lifted (the class layout is unaffected).
#include <cassert>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

// std headers are all included above: lifting the access specifiers only touches the reference header itself
#define private public
#define protected public
#include "core/xs_mesh_homogenized.hpp"
#undef private
#undef protected

#include "core/core_mesh.hpp"
#include "util/error.hpp"

#include "xs_updater.hpp"

using namespace mocc;

namespace mocc_b200 {

namespace {

struct MatTable {          // one material, group-indexed
    std::vector<double> xstr, xsnf, xsf, xsch;
    std::vector<int> row_min, row_max;       // scattering row INTO group ig: source groups [min, max]
    std::vector<std::vector<double>> row;    // row[ig][igg - min]
};

struct PinTable {
    const Pin *pin = nullptr;
    int first_reg  = 0;
    std::vector<int> mat;      // per pin-local FSR: index into the material table
    std::vector<double> area;  // per pin-local FSR
};

struct Homogenizer {
    const XSMeshHomogenized *owner = nullptr;
    int ng                         = 0;
    std::vector<MatTable> mats;
    std::vector<PinTable> pins;

    void build(const XSMeshHomogenized &xs)
    {
        owner = &xs;
        ng    = (int)xs.ng_;
        const MaterialLib &lib = xs.mesh_.mat_lib();
        std::map<int, int> mat_index;
        int reg = 0;
        for (const auto &mplane : xs.mesh_.macroplanes()) {
            for (const auto &pin : mplane) {
                PinTable pt;
                pt.pin                = &*pin;
                pt.first_reg          = reg;
                const auto &pin_mesh  = pin->mesh();
                const VecF &areas     = pin_mesh.areas();
                int ixsreg = 0, local = 0;
                for (const auto &mat_id : pin->mat_ids()) {
                    if (!mat_index.count(mat_id)) {
                        const Material &m = lib.get_material_by_id(mat_id);
                        MatTable t;
                        for (int ig = 0; ig < ng; ig++) {
                            t.xstr.push_back(m.xstr(ig)), t.xsnf.push_back(m.xsnf(ig));
                            t.xsf.push_back(m.xsf(ig)), t.xsch.push_back(m.xsch(ig));
                            const ScatteringRow &r = m.xssc().to(ig);
                            t.row_min.push_back(r.min_g), t.row_max.push_back(r.max_g);
                            t.row.emplace_back(r.from, r.from + (r.max_g - r.min_g + 1));
                        }
                        mat_index[mat_id] = (int)mats.size();
                        mats.push_back(std::move(t));
                    }
                    for (size_t i = 0; i < pin_mesh.n_fsrs(ixsreg); i++, local++) {
                        pt.mat.push_back(mat_index[mat_id]);
                        pt.area.push_back(areas[local]);
                    }
                    ixsreg++;
                }
                reg += pin->n_reg();
                pins.push_back(std::move(pt));
            }
        }
    }

    // homogenize_region_flux (xs_mesh_homogenized.cpp:261-362) for pin i, same operation order
    void pin_update(int i, const ArrayB2 &flux, XSMeshRegion &xsr) const
    {
        const PinTable &pt = pins[i];
        const int nloc     = (int)pt.mat.size();
        std::vector<double> xstr(ng, 0.0), xsnf(ng, 0.0), xsf(ng, 0.0), xsch(ng, 0.0), fs(nloc, 0.0), scatsum(ng);
        std::vector<VecF> scat(ng, VecF(ng, 0.0));
        // fission source per region: the weight of chi (:284-302)
        for (int ig = 0; ig < ng; ig++)
            for (int l = 0; l < nloc; l++)
                fs[l] += mats[pt.mat[l]].xsnf[ig] * flux(pt.first_reg + l, ig) * pt.area[l];
        double fs_sum = 0.0;
        for (const double v : fs)
            fs_sum += v;
        for (int ig = 0; ig < ng; ig++) {
            double fluxvolsum = 0.0;
            std::fill(scatsum.begin(), scatsum.end(), 0.0);
            for (int l = 0; l < nloc; l++) {
                const MatTable &m   = mats[pt.mat[l]];
                const int gmin      = m.row_min[ig], gmax = m.row_max[ig];
                const double v      = pt.area[l];
                const double flux_i = flux(pt.first_reg + l, ig);
                fluxvolsum += v * flux_i;
                xstr[ig] += v * flux_i * m.xstr[ig];
                xsnf[ig] += v * flux_i * m.xsnf[ig];
                xsf[ig] += v * flux_i * m.xsf[ig];
                xsch[ig] += fs[l] * m.xsch[ig];
                for (int igg = 0; igg < ng; igg++) {
                    const double fluxgg = flux(pt.first_reg + l, igg);
                    scatsum[igg] += fluxgg * v;
                    if (igg >= gmin && igg <= gmax)
                        scat[ig][igg] += m.row[ig][igg - gmin] * v * fluxgg;
                }
            }
            for (int igg = 0; igg < ng; igg++)
                if (scat[ig][igg] > 0.0)
                    scat[ig][igg] /= scatsum[igg];
            xstr[ig] /= fluxvolsum;
            xsnf[ig] /= fluxvolsum;
            xsf[ig] /= fluxvolsum;
            if (fs_sum > 0.0)
                xsch[ig] /= fs_sum;
        }
        ScatteringMatrix scat_mat(scat);
        xsr.update(xstr, xsnf, xsch, xsf, scat_mat);
    }
};

bool same_bits(const XSMeshRegion &a, const XSMeshRegion &b, int ng)
{
    for (int ig = 0; ig < ng; ig++) {
        const double x[] = {a.xsmactr(ig), a.xsmacnf(ig), a.xsmacch(ig), a.xsmacf(ig), a.xsmacrm(ig)};
        const double y[] = {b.xsmactr(ig), b.xsmacnf(ig), b.xsmacch(ig), b.xsmacf(ig), b.xsmacrm(ig)};
        if (std::memcmp(x, y, sizeof(x)) != 0)
            return false;
        const ScatteringRow &ra = a.xsmacsc().to(ig), &rb = b.xsmacsc().to(ig);
        if (ra.min_g != rb.min_g || ra.max_g != rb.max_g ||
            std::memcmp(ra.from, rb.from, sizeof(double) * (ra.max_g - ra.min_g + 1)) != 0)
            return false;
    }
    return true;
}

} // namespace

// The flattened tables live as long as the sweeper that owns the updater (which also holds the shared_ptr to the
// mesh they describe): no process-wide state, nothing keyed on an address that could be reused.
struct XsUpdaterImpl {
    Homogenizer hom;
};

XsUpdater::XsUpdater() : impl_(new XsUpdaterImpl())
{
}
XsUpdater::~XsUpdater() = default;

void XsUpdater::update(XSMeshHomogenized &xs)
{
    if (!xs.flux_)
        return; // volume-weighted cross sections: nothing to update (xs_mesh_homogenized.cpp:178-181)
    assert(xs.flux_->extent(0) == (int)xs.mesh_.n_reg(MeshTreatment::PLANE));
    Homogenizer &hom = impl_->hom;
    if (hom.owner != &xs) {
        hom = Homogenizer();
        hom.build(xs);
    }
    static const bool check = std::getenv("MOCB200_CHECK_XS_UPDATE") != nullptr;
    const ArrayB2 &flux = *xs.flux_;
    const int n         = (int)hom.pins.size();
    int bad             = -1;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        hom.pin_update(i, flux, xs.regions_[i]);
        if (check) { // the reference's own routine on a copy of the region: every bit must agree
            XSMeshRegion ref = xs.regions_[i];
            xs.homogenize_region_flux(i, hom.pins[i].first_reg, *hom.pins[i].pin, ref);
            if (!same_bits(ref, xs.regions_[i], hom.ng)) {
#pragma omp critical
                bad = i;
            }
        }
    }
    if (bad >= 0) {
        std::stringstream msg;
        msg << "parallel_update: homogenised cross sections of region " << bad << " differ from the reference routine";
        throw EXCEPT(msg.str());
    }
    xs.state_++;
}

} // namespace mocc_b200
