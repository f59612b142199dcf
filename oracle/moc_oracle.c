/*
 * moc_oracle.c -- plain-C restatement of the reference MoC sweep on the flattened
 * arrays. TEST INFRASTRUCTURE ONLY (see moc_oracle.h). Compiled with
 * -ffp-contract=off so that every operation rounds exactly like the reference
 * build (x86-64, no FMA).
 */
#include "moc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI_ 3.1415926535897932 /* src/core/constants.hpp */
#define FPI_ (4.0 * PI_)

double moc_oracle_exp(const double *d, int n, double vmin, double vmax, double v)
{
    /* exponential.hpp:69-79 */
    double space  = (vmax - vmin) / (double)n;
    double rspace = 1.0 / space;
    if (v < vmin || v > vmax)
        return exp(v);
    int i = (int)((v - vmin) * rspace);
    v -= space * i + vmin;
    return d[i] + (d[i + 1] - d[i]) * v * rspace;
}

void moc_oracle_self_scatter(int n_reg, const double *src, const double *flux, const double *xs_self,
                             const double *xs_tr, double *qbar)
{
    /* source_isotropic.cpp:27-33 */
    for (int i = 0; i < n_reg; i++) {
        double r_fpi_tr = 1.0 / (xs_tr[i] * FPI_);
        qbar[i]         = (src[i] + flux[i] * xs_self[i]) * r_fpi_tr;
    }
}

void moc_oracle_fission_source(int n_reg, int n_group, double k, const double *xs_nf, const double *flux, double *fs)
{
    /* transport_sweeper.cpp:122-131: fission_source = 0; per region, per group, per FSR: += rkeff * xsnf[ig] * flux */
    double rkeff = 1.0 / k;
    for (int r = 0; r < n_reg; r++)
        fs[r] = 0.0;
    for (int g = 0; g < n_group; g++)
        for (int r = 0; r < n_reg; r++)
            fs[r] += rkeff * xs_nf[(size_t)g * n_reg + r] * flux[(size_t)r * n_group + g];
}

void moc_oracle_group_source(int n_reg, int n_group, int group, const double *ext, const double *xs_ch, const double *fs,
                             const double *scat_to, const double *flux, double *src)
{
    for (int r = 0; r < n_reg; r++) {
        double s = ext ? ext[r] : 0.0;  /* source.cpp:43-50 */
        s += xs_ch[r] * fs[r];          /* source.cpp:71-76 */
        for (int gg = 0; gg < n_group; gg++) /* source.cpp:97-108: the row's band in ascending order, self-scatter left out */
            if (gg != group)
                s += scat_to[(size_t)gg * n_reg + r] * flux[(size_t)r * n_group + gg];
        src[r] = s;
    }
}

/* BoundaryCondition::update(group, angle, out), boundary_condition.cpp:155-191 */
static void bc_update_angle(const mocb200_problem *p, int a, double *in, const double *out)
{
    for (int n = 0; n < 2; n++) {
        int size = (n == 0) ? p->bc_size_x[a] : p->bc_size_y[a];
        if (size == 0)
            break;
        int off_out = p->bc_offset[a] + (n == 1 ? p->bc_size_x[a] : 0);
        int off_in  = p->bc_dst_off[2 * a + n];
        switch (p->bc_dst_kind[2 * a + n]) {
        case 0:
            for (int i = 0; i < size; i++)
                in[off_in + i] = 0.0;
            break;
        case 1:
            for (int i = 0; i < size; i++)
                in[off_in + i] = out[off_out + i];
            break;
        default:
            break;
        }
    }
}

static int surf_normal_local(const mocb200_problem *p, int surf_local)
{
    /* Mesh::surface_normal, mesh.hpp:859-871: 2 = Z, 0 = X, 1 = Y */
    if (surf_local < p->nx * p->ny)
        return 2;
    if (surf_local < p->nx * p->ny + (p->nx + 1) * p->ny)
        return 0;
    return 1;
}

/* surface_to_normal for the radial Surface enum values E=0,N=1,W=2,S=3 */
static int surface_to_normal(int s)
{
    return (s == 0 || s == 2) ? 0 : 1;
}

/* calculate_corrections, correction_worker.cpp:32-158, for one (macroplane, sweep angle) */
static void calc_corrections(const mocb200_problem *p, int ip, int a, const double *surf_sum, const double *vol_sum,
                             const double *sigt_sum, const double *sn_xs, double *alpha, double *beta)
{
    /* Surface enum: E=0, N=1, W=2, S=3 */
    int surfs[2][4]; /* [FW/BW][XL, XR, YL, YR] */
    const int n_cell_tot = p->n_plane * p->n_cell_plane;
    surfs[0][2] = 3, surfs[0][3] = 1, surfs[1][2] = 1, surfs[1][3] = 3;
    if (p->ang_ox[a] > 0.0) {
        surfs[0][0] = 2, surfs[0][1] = 0, surfs[1][0] = 0, surfs[1][1] = 2;
    } else {
        surfs[0][0] = 0, surfs[0][1] = 2, surfs[1][0] = 2, surfs[1][1] = 0;
    }
    for (int ic = 0; ic < p->n_cell_plane; ic++) {
        int icc       = ic + p->plane_cell_offset[ip];
        double area_x = p->ang_area_x[a] / p->cell_dx[ic];
        double area_y = p->ang_area_y[a] / p->cell_dy[ic];
        double xstr   = sn_xs[ip * p->n_cell_plane + ic];
        for (int d = 0; d < 2; d++) {
            double psi_xl = surf_sum[p->coarse_surf[4 * ic + surfs[d][0]] * 2 + d] * area_x;
            double psi_xr = surf_sum[p->coarse_surf[4 * ic + surfs[d][1]] * 2 + d] * area_x;
            double psi_yl = surf_sum[p->coarse_surf[4 * ic + surfs[d][2]] * 2 + d] * area_y;
            double psi_yr = surf_sum[p->coarse_surf[4 * ic + surfs[d][3]] * 2 + d] * area_y;
            double ax     = vol_sum[ic * 2 + d] / (psi_xl + psi_xr);
            double ay     = vol_sum[ic * 2 + d] / (psi_yl + psi_yr);
            double b      = sigt_sum[ic * 2 + d] / xstr;
            int iang      = a + d * p->n_ang;
            alpha[((size_t)iang * n_cell_tot + icc) * 2 + 0] = ax;
            alpha[((size_t)iang * n_cell_tot + icc) * 2 + 1] = ay;
            beta[(size_t)iang * n_cell_tot + icc]            = b;
        }
    }
}

static int sweep1g_core(const mocb200_problem *p, int gs_boundary, int tally_mode, const double *xstr,
                        const double *qbar, double *bc_in, double *flux_out, double *current,
                        double *surface_flux, const double *surf_area, const double *xstr_true,
                        const double *sn_xs, double *alpha, double *beta)
{
    const int n_ang = p->n_ang;
    int64_t max_seg = 0;
    for (int64_t t = 0; t < p->n_trk; t++) {
        int64_t n = p->trk_seg_begin[t + 1] - p->trk_seg_begin[t];
        if (n > max_seg)
            max_seg = n;
    }
    double *e_tau  = (double *)malloc(sizeof(double) * (size_t)(max_seg + 1));
    double *psi1   = (double *)malloc(sizeof(double) * (size_t)(max_seg + 1));
    double *psi2   = (double *)malloc(sizeof(double) * (size_t)(max_seg + 1));
    double *bc_out = (double *)calloc((size_t)p->bc_per_group, sizeof(double));
    if (!e_tau || !psi1 || !psi2 || !bc_out)
        return 1;
    /* cmdo::CurrentCorrections per-angle buffers (correction_worker.hpp:49-54) */
    double *surf_sum = NULL, *vol_sum = NULL, *vol_norm = NULL, *sigt_sum = NULL;
    if (tally_mode == 2) {
        surf_sum = (double *)malloc(sizeof(double) * 2 * (size_t)p->n_surf_plane);
        vol_sum  = (double *)malloc(sizeof(double) * 2 * (size_t)p->n_cell_plane);
        sigt_sum = (double *)malloc(sizeof(double) * 2 * (size_t)p->n_cell_plane);
        vol_norm = (double *)malloc(sizeof(double) * (size_t)p->n_cell_plane);
        if (!surf_sum || !vol_sum || !sigt_sum || !vol_norm)
            return 1;
    }

    for (int i = 0; i < p->n_reg; i++)
        flux_out[i] = 0.0;

    for (int ip = 0; ip < p->n_plane; ip++) {
        const int u         = p->plane_unique[ip];
        const int first_reg = p->plane_first_reg[ip];
        double *b_in        = bc_in + (size_t)ip * p->bc_per_group;
        const int cell_off  = p->plane_cell_offset[ip];
        const int surf_off  = p->plane_surf_offset[ip];
        for (int a = 0; a < n_ang; a++) {
            const int a1 = a, a2 = a + n_ang; /* reverse(), angular_quadrature.hpp:158-167 */
            const double *in1 = b_in + p->bc_offset[a1];
            const double *in2 = b_in + p->bc_offset[a2];
            double *out1      = bc_out + p->bc_offset[a1];
            double *out2      = bc_out + p->bc_offset[a2];
            const double rstheta = p->ang_rsintheta[a];
            const double wt_v_st = p->wt_v_st[ip * n_ang + a];
            const double cw[2]   = {p->cur_wx[ip * n_ang + a], p->cur_wy[ip * n_ang + a]};
            const double fw[2]   = {p->flx_wx[ip * n_ang + a], p->flx_wy[ip * n_ang + a]};
            if (tally_mode == 2) { /* set_angle zeroes the sums, correction_worker.hpp:207-221 */
                memset(surf_sum, 0, sizeof(double) * 2 * (size_t)p->n_surf_plane);
                memset(vol_sum, 0, sizeof(double) * 2 * (size_t)p->n_cell_plane);
                memset(sigt_sum, 0, sizeof(double) * 2 * (size_t)p->n_cell_plane);
                memset(vol_norm, 0, sizeof(double) * (size_t)p->n_cell_plane);
            }
            const int geom       = p->ang_geom[a];
            const int64_t t0     = p->geom_trk_begin[(size_t)u * p->n_geom + geom];
            const int64_t t1     = p->geom_trk_begin[(size_t)u * p->n_geom + geom + 1];
            for (int64_t t = t0; t < t1; t++) {
                const int64_t s0 = p->trk_seg_begin[t];
                const int nseg   = (int)(p->trk_seg_begin[t + 1] - s0);
                const double *len = p->seg_len + s0;
                const int32_t *fsr = p->seg_fsr + s0;
                const int bc1 = p->trk_bc[2 * t], bc2 = p->trk_bc[2 * t + 1];

                for (int is = 0; is < nseg; is++) {
                    int ireg  = fsr[is] + first_reg;
                    e_tau[is] = 1.0 - moc_oracle_exp(p->exp_table, p->exp_n, p->exp_min, p->exp_max,
                                                     -xstr[ireg] * len[is] * rstheta);
                }
                /* forward */
                psi1[0] = in1[bc1];
                for (int is = 0; is < nseg; is++) {
                    int ireg        = fsr[is] + first_reg;
                    double psi_diff = (psi1[is] - qbar[ireg]) * e_tau[is];
                    psi1[is + 1]    = psi1[is] - psi_diff;
                    flux_out[ireg] += psi_diff * wt_v_st;
                }
                out1[bc2] = psi1[nseg];
                /* backward */
                psi2[nseg] = in2[bc2];
                for (int is = nseg - 1; is >= 0; is--) {
                    int ireg        = fsr[is] + first_reg;
                    double psi_diff = (psi2[is + 1] - qbar[ireg]) * e_tau[is];
                    psi2[is]        = psi2[is + 1] - psi_diff;
                    flux_out[ireg] += psi_diff * wt_v_st;
                }
                out2[bc1] = psi2[0];

                if (tally_mode == 1) {
                    /* moc::Current::post_ray, moc_current_worker.hpp:202-264 */
                    int cell_fw = p->trk_cm_start[4 * t + 0] + cell_off;
                    int cell_bw = p->trk_cm_start[4 * t + 1] + cell_off;
                    int surf_fw = p->trk_cm_start[4 * t + 2] + surf_off;
                    int surf_bw = p->trk_cm_start[4 * t + 3] + surf_off;
                    int iseg_fw = 0, iseg_bw = nseg;
                    int norm_fw = surf_normal_local(p, surf_fw - surf_off);
                    int norm_bw = surf_normal_local(p, surf_bw - surf_off);
                    current[surf_fw] += psi1[iseg_fw] * cw[norm_fw];
                    current[surf_bw] -= psi2[iseg_bw] * cw[norm_bw];
                    surface_flux[surf_fw] += psi1[iseg_fw] * fw[norm_fw];
                    surface_flux[surf_bw] += psi2[iseg_bw] * fw[norm_bw];
                    for (int64_t k = p->trk_cm_begin[t]; k < p->trk_cm_begin[t + 1]; k++) {
                        uint32_t c = p->cm_data[k];
                        int s_fw = c & 0xF, s_bw = (c >> 4) & 0xF;
                        int n_fw = (c >> 8) & 0xFF, n_bw = (c >> 16) & 0xFF;
                        if (s_fw != 7) {
                            iseg_fw += n_fw;
                            norm_fw = surface_to_normal(s_fw);
                            surf_fw = p->coarse_surf[4 * (cell_fw - cell_off) + s_fw] + surf_off;
                            current[surf_fw] += psi1[iseg_fw] * cw[norm_fw];
                            surface_flux[surf_fw] += psi1[iseg_fw] * fw[norm_fw];
                        }
                        if (s_bw != 7) {
                            iseg_bw -= n_bw;
                            norm_bw = surface_to_normal(s_bw);
                            surf_bw = p->coarse_surf[4 * (cell_bw - cell_off) + s_bw] + surf_off;
                            current[surf_bw] -= psi2[iseg_bw] * cw[norm_bw];
                            surface_flux[surf_bw] += psi2[iseg_bw] * fw[norm_bw];
                        }
                        /* coarse_neighbor(cell, INVALID) returns -5 in the reference and the cell
                         * is never used again on that side; keep the cell instead */
                        if (s_fw < 4) {
                            int nb  = p->coarse_nbr[4 * (cell_fw - cell_off) + s_fw];
                            cell_fw = (nb < 0) ? cell_fw : nb + cell_off;
                        }
                        if (s_bw < 4) {
                            int nb  = p->coarse_nbr[4 * (cell_bw - cell_off) + s_bw];
                            cell_bw = (nb < 0) ? cell_bw : nb + cell_off;
                        }
                    }
                }
                if (tally_mode == 2) {
                    /* cmdo::CurrentCorrections::post_ray, correction_worker.hpp:109-205: currents use the
                     * plane offset, the per-angle sums are plane-local; the backward SURFACE FLUX is
                     * subtracted here (:136-137, :194-195) where moc::Current adds it */
                    int cell_fw = p->trk_cm_start[4 * t + 0], cell_bw = p->trk_cm_start[4 * t + 1];
                    int surf_fw = p->trk_cm_start[4 * t + 2], surf_bw = p->trk_cm_start[4 * t + 3];
                    int iseg_fw = 0, iseg_bw = nseg;
                    int norm_fw = surf_normal_local(p, surf_fw), norm_bw = surf_normal_local(p, surf_bw);
                    current[surf_fw + surf_off] += psi1[iseg_fw] * cw[norm_fw];
                    current[surf_bw + surf_off] -= psi2[iseg_bw] * cw[norm_bw];
                    surface_flux[surf_fw + surf_off] += psi1[iseg_fw] * fw[norm_fw];
                    surface_flux[surf_bw + surf_off] -= psi2[iseg_bw] * fw[norm_bw];
                    surf_sum[surf_fw * 2 + 0] += psi1[iseg_fw];
                    surf_sum[surf_bw * 2 + 1] += psi2[iseg_bw];
                    for (int64_t k = p->trk_cm_begin[t]; k < p->trk_cm_begin[t + 1]; k++) {
                        uint32_t c = p->cm_data[k];
                        int s_fw = c & 0xF, s_bw = (c >> 4) & 0xF;
                        int n_fw = (c >> 8) & 0xFF, n_bw = (c >> 16) & 0xFF;
                        if (s_fw != 7) {
                            for (int i = 0; i < n_fw; i++) {
                                int ireg       = fsr[iseg_fw] + first_reg;
                                double tt      = rstheta * len[iseg_fw];
                                double fluxvol = tt * qbar[ireg] + (psi1[iseg_fw] - psi1[iseg_fw + 1]) / xstr[ireg];
                                vol_sum[cell_fw * 2 + 0] += fluxvol;
                                vol_norm[cell_fw] += tt;
                                sigt_sum[cell_fw * 2 + 0] += xstr_true[ireg] * fluxvol;
                                iseg_fw++;
                            }
                            norm_fw = surface_to_normal(s_fw);
                            surf_fw = p->coarse_surf[4 * cell_fw + s_fw];
                            current[surf_fw + surf_off] += psi1[iseg_fw] * cw[norm_fw];
                            surface_flux[surf_fw + surf_off] += psi1[iseg_fw] * fw[norm_fw];
                            surf_sum[surf_fw * 2 + 0] += psi1[iseg_fw];
                        }
                        if (s_bw != 7) {
                            for (int i = 0; i < n_bw; i++) {
                                iseg_bw--;
                                int ireg       = fsr[iseg_bw] + first_reg;
                                double tt      = rstheta * len[iseg_bw];
                                double fluxvol = tt * qbar[ireg] +
                                                 e_tau[iseg_bw] * (psi2[iseg_bw + 1] - qbar[ireg]) / xstr[ireg];
                                vol_sum[cell_bw * 2 + 1] += fluxvol;
                                sigt_sum[cell_bw * 2 + 1] += xstr_true[ireg] * fluxvol;
                            }
                            norm_bw = surface_to_normal(s_bw);
                            surf_bw = p->coarse_surf[4 * cell_bw + s_bw];
                            current[surf_bw + surf_off] -= psi2[iseg_bw] * cw[norm_bw];
                            surface_flux[surf_bw + surf_off] -= psi2[iseg_bw] * fw[norm_bw];
                            surf_sum[surf_bw * 2 + 1] += psi2[iseg_bw];
                        }
                        if (s_fw < 4) {
                            int nb  = p->coarse_nbr[4 * cell_fw + s_fw];
                            cell_fw = (nb < 0) ? cell_fw : nb;
                        }
                        if (s_bw < 4) {
                            int nb  = p->coarse_nbr[4 * cell_bw + s_bw];
                            cell_bw = (nb < 0) ? cell_bw : nb;
                        }
                    }
                }
            } /* rays */
            if (tally_mode == 2) { /* post_angle, correction_worker.hpp:223-246 */
                for (int i = 0; i < p->n_cell_plane; i++) {
                    sigt_sum[2 * i + 0] /= vol_sum[2 * i + 0];
                    sigt_sum[2 * i + 1] /= vol_sum[2 * i + 1];
                    vol_sum[2 * i + 0] /= vol_norm[i];
                    vol_sum[2 * i + 1] /= vol_norm[i];
                }
                calc_corrections(p, ip, a, surf_sum, vol_sum, sigt_sum, sn_xs, alpha, beta);
            }
            if (gs_boundary) {
                bc_update_angle(p, a1, b_in, bc_out);
                bc_update_angle(p, a2, b_in, bc_out);
            }
        } /* angles */
        if (!gs_boundary) {
            for (int a = 0; a < 2 * n_ang; a++)
                bc_update_angle(p, a, b_in, bc_out);
        }
    } /* planes */

    /* kernel:165-173 */
    for (int i = 0; i < p->n_reg; i++)
        flux_out[i] = flux_out[i] / (xstr[i] * p->vol[i]) + qbar[i] * FPI_;

    if (tally_mode == 2) {
        free(surf_sum);
        free(vol_sum);
        free(sigt_sum);
        free(vol_norm);
    }
    if (tally_mode >= 1) {
        /* post_sweep normalisation, moc_current_worker.hpp:303-316 (no sub-plane expansion here:
         * callers with sub-planes expand on the host exactly like the reference) */
        for (int ip = 0; ip < p->n_plane; ip++) {
            int b = p->plane_surf_offset[ip] + p->nx * p->ny;
            int e = p->plane_surf_offset[ip] + p->n_surf_plane;
            for (int s = b; s < e; s++) {
                current[s] /= surf_area[s];
                surface_flux[s] /= surf_area[s];
            }
        }
    }
    free(e_tau);
    free(psi1);
    free(psi2);
    free(bc_out);
    return 0;
}

int moc_oracle_sweep1g(const mocb200_problem *p, int gs_boundary, int tally_mode, const double *xstr,
                       const double *qbar, double *bc_in, double *flux_out, double *current,
                       double *surface_flux, const double *surf_area)
{
    if (tally_mode != 0 && tally_mode != 1)
        return 2;
    return sweep1g_core(p, gs_boundary, tally_mode, xstr, qbar, bc_in, flux_out, current, surface_flux, surf_area,
                        NULL, NULL, NULL, NULL);
}

int moc_oracle_sweep1g_corrections(const mocb200_problem *p, int gs_boundary, const double *xstr_split,
                                   const double *xstr_true, const double *qbar, const double *sn_xs, double *bc_in,
                                   double *flux_out, double *current, double *surface_flux, const double *surf_area,
                                   double *alpha, double *beta)
{
    if (!p->ang_area_x || !p->ang_area_y || !p->ang_ox || !p->cell_dx || !p->cell_dy)
        return 2;
    return sweep1g_core(p, gs_boundary, 2, xstr_split, qbar, bc_in, flux_out, current, surface_flux, surf_area,
                        xstr_true, sn_xs, alpha, beta);
}
