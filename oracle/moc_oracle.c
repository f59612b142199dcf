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

int moc_oracle_sweep1g(const mocb200_problem *p, int gs_boundary, int tally_mode, const double *xstr,
                       const double *qbar, double *bc_in, double *flux_out, double *current,
                       double *surface_flux, const double *surf_area)
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
            } /* rays */
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

    if (tally_mode == 1) {
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
