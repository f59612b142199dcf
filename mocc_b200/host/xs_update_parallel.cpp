// xs_update_parallel.cpp -- XSMeshHomogenized::update() (src/core/xs_mesh_homogenized.cpp:176-197) with the
// pin loop spread over the host threads.
//
// MoCSweeper_2D3D::sweep re-homogenises the Sn cross-section mesh before the last inner of EVERY group
// (moc_sweeper_2d3d.cpp:85-88). The reference does that serially, one pin after the other, and on C5G7-sized
// planes it is what a sweep(group) call costs once the MoC sweep itself runs on the GPU (profiles/r1: 0.48 s of
// 0.54 s per call on the 9-plane C5G7 3-D case). Every pin writes only its own XSMeshRegion and reads shared
// state that is constant during the update, so the pins are independent: this file calls the reference's OWN
// per-pin routine (homogenize_region_flux, unchanged arithmetic, bit-identical results) from an OpenMP loop.
// The routine is private to the class; this translation unit alone sees the reference headers with that
// access lifted (the class layout is unaffected).
#include <cassert>
#include <cstddef>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#define private public
#define protected public
#include "core/xs_mesh_homogenized.hpp"
#undef private
#undef protected

#include "core/core_mesh.hpp"

namespace mocc_b200 {

void parallel_update(mocc::XSMeshHomogenized &xs)
{
    if (!xs.flux_)
        return; // volume-weighted cross sections: nothing to update (xs_mesh_homogenized.cpp:178-181)
    assert(xs.flux_->extent(0) == (int)xs.mesh_.n_reg(mocc::MeshTreatment::PLANE));
    std::vector<const mocc::Pin *> pins;
    std::vector<int> first_reg;
    int reg = 0;
    for (const auto &mplane : xs.mesh_.macroplanes()) {
        for (const auto &pin : mplane) {
            pins.push_back(&*pin);
            first_reg.push_back(reg);
            reg += pin->n_reg();
        }
    }
    const int n = (int)pins.size();
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++)
        xs.homogenize_region_flux(i, first_reg[i], *pins[i], xs.regions_[i]);
    xs.state_++;
}

} // namespace mocc_b200
