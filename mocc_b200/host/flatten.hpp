// flatten.hpp -- one-time flattening of MOCC's ray-tracing data into the SoA
// arrays the B200 sweep consumes (include/mocc_b200.h: mocb200_problem).
//
// Everything here is READ from reference objects; nothing about the geometry is
// re-derived, so FSR indexing, boundary linkage and coarse-ray linkage are
// bit-exact by construction:
//   rays / segments / bc / cm_data   moc::Ray      sweepers/moc/ray.hpp:33-216
//   Nx, Ny, spacing, modularised quadrature  moc::RayData  sweepers/moc/ray_data.hpp:83-227
//   boundary layout  BoundaryCondition  core/boundary_condition.cpp:33-76, 145-191
//   coarse surfaces  Mesh  core/mesh.hpp:482-487, 779-822
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "arrayfile.hpp"
#include "mocc_b200.h"

namespace mocc {
class CoreMesh;
class AngularQuadrature;
namespace moc {
class RayData;
class Ray;
}
}

namespace mocc_b200 {

struct FlatProblem {
    // scalars
    int32_t n_group = 0, n_reg = 0, n_plane = 0, n_unique = 0, ndir_oct = 0, n_ang = 0, n_geom = 0;
    int32_t bc_per_group = 0, n_surf = 0, n_cell = 0, n_surf_plane = 0, n_cell_plane = 0;
    int32_t nx = 0, ny = 0, nz = 0, exp_n = 10000;
    double exp_min = -10.0, exp_max = 0.0;

    std::vector<int32_t> ang_geom;
    std::vector<double> ang_rsintheta;
    // informational per-angle data (not consumed by the device)
    std::vector<double> ang_alpha, ang_theta, ang_weight, ang_spacing;
    std::vector<double> wt_v_st, cur_wx, cur_wy, flx_wx, flx_wy;
    std::vector<int32_t> bc_offset, bc_size_x, bc_size_y, bc_dst_off, bc_dst_kind;
    std::vector<int64_t> geom_trk_begin, trk_seg_begin, trk_cm_begin;
    std::vector<int32_t> trk_bc, trk_cm_start, seg_fsr;
    std::vector<double> seg_len;
    std::vector<uint32_t> cm_data;
    std::vector<int32_t> plane_unique, plane_first_reg, plane_cell_offset, plane_surf_offset;
    std::vector<double> plane_height, plane_dz;
    std::vector<int32_t> coarse_surf, coarse_nbr;
    std::vector<double> vol, surf_area, exp_table;
    // 2D3D correction-factor tables (cmdo::CurrentCorrections)
    std::vector<double> ang_area_x, ang_area_y, ang_ox, cell_dx, cell_dy;
    std::vector<int32_t> plane_xs_offset; // CurrentCorrections::mplane_offset_ (index into PIN-expanded XS)
    // reference segment count S: sum over macroplanes and ALL sweep angles (polar copies counted)
    int64_t n_seg_reference = 0;
    int64_t n_ray_reference = 0;

    mocb200_problem view() const;
    ArrayFile to_arrayfile() const;
    static FlatProblem from_arrayfile(const ArrayFile &af);
};

// Appends one reference Ray as one track (what flatten() does for every ray of every geometry class)
void append_ray(FlatProblem &fp, const mocc::moc::Ray &ray);

// vol: FSR volumes as TransportSweeper::vol_ holds them (MeshTreatment::PLANE)
FlatProblem flatten(const mocc::CoreMesh &mesh, const mocc::moc::RayData &rays,
                    const std::vector<int> &macroplane_unique_ids,
                    const std::vector<int> &first_reg_macroplane, const double *vol,
                    int n_reg, int n_group);
}
