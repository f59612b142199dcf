#include "flatten.hpp"

#include <cmath>
#include <cstring>
#include <stdexcept>

#include "core/angular_quadrature.hpp"
#include "core/constants.hpp"
#include "core/core_mesh.hpp"
#include "ray_data.hpp"

namespace mocc_b200 {

namespace {
using mocc::moc::Ray;

uint32_t pack_cm(const Ray &ray, size_t i)
{
    const auto &c = ray.cm_data()[i];
    return (uint32_t)c.fw | ((uint32_t)c.bw << 4) | ((uint32_t)c.nseg_fw << 8) |
           ((uint32_t)c.nseg_bw << 16);
}

// Two rays carry the same data, bit for bit
bool same_ray(const Ray &a, const Ray &b)
{
    if (a.nseg() != b.nseg() || a.ncseg() != b.ncseg())
        return false;
    if (a.bc(0) != b.bc(0) || a.bc(1) != b.bc(1))
        return false;
    if (a.cm_cell_fw() != b.cm_cell_fw() || a.cm_cell_bw() != b.cm_cell_bw() ||
        a.cm_surf_fw() != b.cm_surf_fw() || a.cm_surf_bw() != b.cm_surf_bw())
        return false;
    if (a.nseg() > 0) {
        if (std::memcmp(a.seg_len().data(), b.seg_len().data(), sizeof(double) * a.nseg()) != 0)
            return false;
        if (std::memcmp(a.seg_index().data(), b.seg_index().data(), sizeof(int) * a.nseg()) != 0)
            return false;
    }
    for (int i = 0; i < a.ncseg(); i++)
        if (pack_cm(a, i) != pack_cm(b, i))
            return false;
    return true;
}
}

// One reference Ray (ray.hpp:33-216) -> one track of the flat arrays: boundary slots, the coarse cells / surfaces
// it starts from in either direction, segment lengths and FSR ids in forward order, packed RayCoarseData records.
void append_ray(FlatProblem &fp, const mocc::moc::Ray &ray)
{
    if (fp.trk_seg_begin.empty())
        fp.trk_seg_begin.push_back(0);
    if (fp.trk_cm_begin.empty())
        fp.trk_cm_begin.push_back(0);
    fp.trk_bc.push_back(ray.bc(0));
    fp.trk_bc.push_back(ray.bc(1));
    fp.trk_cm_start.push_back((int)ray.cm_cell_fw());
    fp.trk_cm_start.push_back((int)ray.cm_cell_bw());
    fp.trk_cm_start.push_back((int)ray.cm_surf_fw());
    fp.trk_cm_start.push_back((int)ray.cm_surf_bw());
    for (int is = 0; is < ray.nseg(); is++) {
        fp.seg_len.push_back(ray.seg_len(is));
        fp.seg_fsr.push_back((int)ray.seg_index(is));
    }
    for (int ic = 0; ic < ray.ncseg(); ic++)
        fp.cm_data.push_back(pack_cm(ray, ic));
    fp.trk_seg_begin.push_back((int64_t)fp.seg_len.size());
    fp.trk_cm_begin.push_back((int64_t)fp.cm_data.size());
}

FlatProblem flatten(const mocc::CoreMesh &mesh, const mocc::moc::RayData &rays,
                    const std::vector<int> &macroplane_unique_ids,
                    const std::vector<int> &first_reg_macroplane, const double *vol,
                    int n_reg, int n_group)
{
    using namespace mocc;
    FlatProblem fp;
    const AngularQuadrature &aq = rays.ang_quad();

    fp.n_group      = n_group;
    fp.n_reg        = n_reg;
    fp.n_plane      = (int)macroplane_unique_ids.size();
    fp.n_unique     = (int)std::distance(rays.begin(), rays.end());
    fp.ndir_oct     = aq.ndir_oct();
    fp.n_ang        = 2 * fp.ndir_oct;
    fp.nx           = (int)mesh.nx();
    fp.ny           = (int)mesh.ny();
    fp.nz           = (int)mesh.nz();
    fp.n_cell_plane = fp.nx * fp.ny;
    fp.n_surf_plane = fp.nx * fp.ny + (fp.nx + 1) * fp.ny + (fp.ny + 1) * fp.nx;
    fp.n_surf       = (int)mesh.n_surf();
    fp.n_cell       = (int)mesh.n_pin();
    if (mesh.coarse_surf_offset(1) != fp.n_surf_plane || mesh.coarse_cell_offset(1) != fp.n_cell_plane)
        throw std::runtime_error("flatten: unexpected coarse-mesh plane stride");

    const int n_ang    = fp.n_ang;
    const int n_ang_bc = 2 * n_ang;
    if ((int)rays.begin()->size() != n_ang)
        throw std::runtime_error("flatten: ray set does not span octants 1-2");

    // ---- angles: geometry classes (polar copies of one azimuth are bit-identical) ----
    fp.ang_geom.assign(n_ang, -1);
    std::vector<int> geom_first; // representative angle of each geometry
    for (int a = 0; a < n_ang; a++) {
        for (size_t gi = 0; gi < geom_first.size() && fp.ang_geom[a] < 0; gi++) {
            int b = geom_first[gi];
            if (rays.nx(a) != rays.nx(b) || rays.ny(a) != rays.ny(b))
                continue;
            bool same = true;
            for (int u = 0; u < fp.n_unique && same; u++) {
                const auto &ra = rays[u][a];
                const auto &rb = rays[u][b];
                same = ra.size() == rb.size();
                for (size_t i = 0; i < ra.size() && same; i++)
                    same = same_ray(ra[i], rb[i]);
            }
            if (same)
                fp.ang_geom[a] = (int)gi;
        }
        if (fp.ang_geom[a] < 0) {
            fp.ang_geom[a] = (int)geom_first.size();
            geom_first.push_back(a);
        }
    }
    fp.n_geom = (int)geom_first.size();

    for (int a = 0; a < n_ang; a++) {
        const Angle &ang = aq[a];
        fp.ang_rsintheta.push_back(ang.rsintheta);
        fp.ang_alpha.push_back(ang.alpha);
        fp.ang_theta.push_back(ang.theta);
        fp.ang_weight.push_back(ang.weight);
        fp.ang_spacing.push_back(rays.spacing(a));
    }

    // ---- per (macroplane, angle) weights; same operation order as the reference ----
    for (int ip = 0; ip < fp.n_plane; ip++) {
        const real_t height = mesh.macroplanes()[ip].height;
        // quirk kept: Current::set_angle indexes dz with the MACROplane index
        // (moc_current_worker.hpp:190)
        const real_t dz = mesh.dz(ip);
        fp.plane_height.push_back(height);
        fp.plane_dz.push_back(dz);
        for (int a = 0; a < n_ang; a++) {
            const Angle &ang     = aq[a];
            const real_t spacing = rays.spacing(a);
            real_t stheta        = std::sin(ang.theta);
            fp.wt_v_st.push_back(ang.weight * spacing * height * stheta * PI);
            real_t w = ang.weight * PI;
            fp.cur_wx.push_back(w * ang.ox * spacing / std::abs(std::cos(ang.alpha)) * dz);
            fp.cur_wy.push_back(w * ang.oy * spacing / std::abs(std::sin(ang.alpha)) * dz);
            fp.flx_wx.push_back(w * spacing / std::abs(std::cos(ang.alpha)) * dz);
            fp.flx_wy.push_back(w * spacing / std::abs(std::sin(ang.alpha)) * dz);
        }
    }

    // ---- 2D3D correction-factor tables, same expressions as correction_worker.cpp:78-93 ----
    for (int a = 0; a < n_ang; a++) {
        fp.ang_area_x.push_back(std::abs(rays.spacing(a) / cos(aq[a].alpha)));
        fp.ang_area_y.push_back(std::abs(rays.spacing(a) / sin(aq[a].alpha)));
        fp.ang_ox.push_back(aq[a].ox);
    }
    for (int c = 0; c < fp.n_cell_plane; c++) {
        auto pos = mesh.coarse_position(c);
        fp.cell_dx.push_back(mesh.pin_dx()[pos.x]);
        fp.cell_dy.push_back(mesh.pin_dy()[pos.y]);
    }
    {
        // CurrentCorrections::mplane_offset_ (correction_worker.hpp:66-78)
        int mplane = 0, offset = 0;
        fp.plane_xs_offset.push_back(0);
        for (const auto index : mesh.macroplane_index()) {
            if (index != mplane) {
                fp.plane_xs_offset.push_back(offset);
                mplane = index;
            }
            offset += fp.n_cell_plane;
        }
    }

    // ---- boundary layout (BoundaryCondition ctor + update) ----
    fp.bc_size_x.assign(n_ang_bc, 0);
    fp.bc_size_y.assign(n_ang_bc, 0);
    for (int a = 0; a < n_ang; a++) {
        int r           = aq.reverse(a);
        fp.bc_size_x[a] = fp.bc_size_x[r] = (int)rays.ny(a);
        fp.bc_size_y[a] = fp.bc_size_y[r] = (int)rays.nx(a);
    }
    fp.bc_offset.assign(n_ang_bc, 0);
    int off = 0;
    for (int a = 0; a < n_ang_bc; a++) {
        fp.bc_offset[a] = off;
        off += fp.bc_size_x[a] + fp.bc_size_y[a];
    }
    fp.bc_per_group = off;
    fp.bc_dst_off.assign(2 * n_ang_bc, 0);
    fp.bc_dst_kind.assign(2 * n_ang_bc, 2);
    const auto &bc_type = mesh.boundary();
    for (int a = 0; a < n_ang_bc; a++) {
        for (int n = 0; n < 2; n++) {
            Normal norm  = (n == 0) ? Normal::X_NORM : Normal::Y_NORM;
            int a_in     = aq.reflect(a, norm);
            int face_off = fp.bc_offset[a_in] + (n == 1 ? fp.bc_size_x[a_in] : 0);
            if ((n == 0 ? fp.bc_size_x[a_in] : fp.bc_size_y[a_in]) !=
                (n == 0 ? fp.bc_size_x[a] : fp.bc_size_y[a]))
                throw std::runtime_error("flatten: reflected boundary face size mismatch");
            fp.bc_dst_off[2 * a + n] = face_off;
            int kind;
            switch (bc_type[(int)(aq[a_in].upwind_surface(norm))]) {
            case Boundary::VACUUM:
                kind = 0;
                break;
            case Boundary::REFLECT:
                kind = 1;
                break;
            case Boundary::PRESCRIBED:
                kind = 2;
                break;
            default:
                throw std::runtime_error("flatten: unsupported boundary condition type");
            }
            fp.bc_dst_kind[2 * a + n] = kind;
        }
    }

    // ---- tracks / segments / coarse-ray records, one copy per geometry ----
    fp.geom_trk_begin.push_back(0);
    fp.trk_seg_begin.push_back(0);
    fp.trk_cm_begin.push_back(0);
    for (int u = 0; u < fp.n_unique; u++) {
        for (int gi = 0; gi < fp.n_geom; gi++) {
            const auto &ang_rays = rays[u][geom_first[gi]];
            for (const auto &ray : ang_rays)
                append_ray(fp, ray);
            fp.geom_trk_begin.push_back((int64_t)fp.trk_bc.size() / 2);
        }
    }

    // ---- macroplanes ----
    for (int ip = 0; ip < fp.n_plane; ip++) {
        fp.plane_unique.push_back(macroplane_unique_ids[ip]);
        fp.plane_first_reg.push_back(first_reg_macroplane[ip]);
        fp.plane_cell_offset.push_back(mesh.coarse_cell_offset(ip));
        fp.plane_surf_offset.push_back(mesh.coarse_surf_offset(ip));
        int u = macroplane_unique_ids[ip];
        for (int a = 0; a < n_ang; a++) {
            fp.n_ray_reference += (int64_t)rays[u][a].size();
            for (const auto &ray : rays[u][a])
                fp.n_seg_reference += ray.nseg();
        }
    }

    // ---- coarse mesh connectivity of one plane ----
    const Surface radial[4] = {Surface::EAST, Surface::NORTH, Surface::WEST, Surface::SOUTH};
    for (int c = 0; c < fp.n_cell_plane; c++) {
        for (int s = 0; s < 4; s++) {
            fp.coarse_surf.push_back(mesh.coarse_surf(c, radial[s]));
            fp.coarse_nbr.push_back(mesh.coarse_neighbor(c, radial[s]));
        }
    }
    fp.surf_area.resize(fp.n_surf);
    for (int s = 0; s < fp.n_surf; s++)
        fp.surf_area[s] = mesh.coarse_area(s);

    fp.vol.assign(vol, vol + n_reg);

    // ---- exponential table, as Exponential_Linear<N>'s constructor builds it ----
    {
        real_t space = (fp.exp_max - fp.exp_min) / (real_t)(fp.exp_n);
        fp.exp_table.resize(fp.exp_n + 2);
        for (int i = 0; i <= fp.exp_n; i++)
            fp.exp_table[i] = std::exp(fp.exp_min + i * space);
        fp.exp_table[fp.exp_n + 1] = fp.exp_table[fp.exp_n];
    }
    return fp;
}

mocb200_problem FlatProblem::view() const
{
    mocb200_problem p;
    std::memset(&p, 0, sizeof(p));
    p.n_group = n_group, p.n_reg = n_reg, p.n_plane = n_plane, p.n_unique = n_unique;
    p.ndir_oct = ndir_oct, p.n_ang = n_ang, p.n_geom = n_geom, p.bc_per_group = bc_per_group;
    p.n_surf = n_surf, p.n_cell = n_cell, p.n_surf_plane = n_surf_plane, p.n_cell_plane = n_cell_plane;
    p.nx = nx, p.ny = ny, p.nz = nz, p.exp_n = exp_n, p.exp_min = exp_min, p.exp_max = exp_max;
    p.n_trk = (int64_t)trk_bc.size() / 2;
    p.n_seg = (int64_t)seg_len.size();
    p.n_cm  = (int64_t)cm_data.size();
    p.ang_geom = ang_geom.data(), p.ang_rsintheta = ang_rsintheta.data();
    p.wt_v_st = wt_v_st.data(), p.cur_wx = cur_wx.data(), p.cur_wy = cur_wy.data();
    p.flx_wx = flx_wx.data(), p.flx_wy = flx_wy.data();
    p.bc_offset = bc_offset.data(), p.bc_size_x = bc_size_x.data(), p.bc_size_y = bc_size_y.data();
    p.bc_dst_off = bc_dst_off.data(), p.bc_dst_kind = bc_dst_kind.data();
    p.geom_trk_begin = geom_trk_begin.data(), p.trk_seg_begin = trk_seg_begin.data();
    p.trk_bc = trk_bc.data(), p.trk_cm_begin = trk_cm_begin.data(), p.trk_cm_start = trk_cm_start.data();
    p.seg_len = seg_len.data(), p.seg_fsr = seg_fsr.data(), p.cm_data = cm_data.data();
    p.plane_unique = plane_unique.data(), p.plane_first_reg = plane_first_reg.data();
    p.plane_cell_offset = plane_cell_offset.data(), p.plane_surf_offset = plane_surf_offset.data();
    p.coarse_surf = coarse_surf.data(), p.coarse_nbr = coarse_nbr.data();
    p.vol = vol.data(), p.exp_table = exp_table.data();
    p.ang_area_x = ang_area_x.data(), p.ang_area_y = ang_area_y.data(), p.ang_ox = ang_ox.data();
    p.cell_dx = cell_dx.data(), p.cell_dy = cell_dy.data();
    return p;
}

ArrayFile FlatProblem::to_arrayfile() const
{
    ArrayFile af;
#define S_(x) af.put_scalar(#x, x)
#define V_(x) af.put(#x, x)
    S_(n_group), S_(n_reg), S_(n_plane), S_(n_unique), S_(ndir_oct), S_(n_ang), S_(n_geom);
    S_(bc_per_group), S_(n_surf), S_(n_cell), S_(n_surf_plane), S_(n_cell_plane);
    S_(nx), S_(ny), S_(nz), S_(exp_n), S_(exp_min), S_(exp_max);
    S_(n_seg_reference), S_(n_ray_reference);
    V_(ang_geom), V_(ang_rsintheta), V_(ang_alpha), V_(ang_theta), V_(ang_weight), V_(ang_spacing);
    V_(wt_v_st), V_(cur_wx), V_(cur_wy), V_(flx_wx), V_(flx_wy);
    V_(bc_offset), V_(bc_size_x), V_(bc_size_y), V_(bc_dst_off), V_(bc_dst_kind);
    V_(geom_trk_begin), V_(trk_seg_begin), V_(trk_bc), V_(trk_cm_begin), V_(trk_cm_start);
    V_(seg_len), V_(seg_fsr), V_(cm_data);
    V_(plane_unique), V_(plane_first_reg), V_(plane_cell_offset), V_(plane_surf_offset);
    V_(plane_height), V_(plane_dz), V_(coarse_surf), V_(coarse_nbr), V_(vol), V_(surf_area), V_(exp_table);
    V_(ang_area_x), V_(ang_area_y), V_(ang_ox), V_(cell_dx), V_(cell_dy), V_(plane_xs_offset);
#undef S_
#undef V_
    return af;
}

namespace {
template <class T> void take(const ArrayFile &af, const char *name, std::vector<T> &v)
{
    const NamedArray &a = af.get(name);
    const T *p          = a.as<T>();
    v.assign(p, p + a.count());
}
}

FlatProblem FlatProblem::from_arrayfile(const ArrayFile &af)
{
    FlatProblem fp;
#define S_(x) fp.x = af.scalar<decltype(fp.x)>(#x)
#define V_(x) take(af, #x, fp.x)
    S_(n_group), S_(n_reg), S_(n_plane), S_(n_unique), S_(ndir_oct), S_(n_ang), S_(n_geom);
    S_(bc_per_group), S_(n_surf), S_(n_cell), S_(n_surf_plane), S_(n_cell_plane);
    S_(nx), S_(ny), S_(nz), S_(exp_n), S_(exp_min), S_(exp_max);
    S_(n_seg_reference), S_(n_ray_reference);
    V_(ang_geom), V_(ang_rsintheta), V_(ang_alpha), V_(ang_theta), V_(ang_weight), V_(ang_spacing);
    V_(wt_v_st), V_(cur_wx), V_(cur_wy), V_(flx_wx), V_(flx_wy);
    V_(bc_offset), V_(bc_size_x), V_(bc_size_y), V_(bc_dst_off), V_(bc_dst_kind);
    V_(geom_trk_begin), V_(trk_seg_begin), V_(trk_bc), V_(trk_cm_begin), V_(trk_cm_start);
    V_(seg_len), V_(seg_fsr), V_(cm_data);
    V_(plane_unique), V_(plane_first_reg), V_(plane_cell_offset), V_(plane_surf_offset);
    V_(plane_height), V_(plane_dz), V_(coarse_surf), V_(coarse_nbr), V_(vol), V_(surf_area), V_(exp_table);
    V_(ang_area_x), V_(ang_area_y), V_(ang_ox), V_(cell_dx), V_(cell_dy), V_(plane_xs_offset);
#undef S_
#undef V_
    return fp;
}
}
