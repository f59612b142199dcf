/*
 * mocc_b200.h -- C ABI of the B200-native MoC transport sweep.
 *
 * This is the drop-in boundary: the only thing the host
 * side (the C++ TransportSweeper subclass in mocc_b200/host/, or any FFI such as
 * the ctypes binding in mocc_b200/capi.py) ever calls. Plain pointers and sizes,
 * caller-owned host buffers, callee-owned device memory, int status codes, no
 * exceptions and no torch/C++ types across the boundary.
 *
 * Reference interfaces replaced (all under /root/reference/src):
 *   mocb200_create            <- moc::MoCSweeper::MoCSweeper          sweepers/moc/moc_sweeper.cpp:62-187
 *                                (device copy of RayData / BoundaryCondition layout / Exponential_Linear)
 *   mocb200_set_xs            <- ExpandedXS::expand                   core/xs_mesh.hpp:261-298
 *   mocb200_set_source        <- Source::get() handed to self_scatter core/source.hpp:150-153
 *   mocb200_set_flux/get_flux <- TransportSweeper::flux_ column       core/transport_sweeper.hpp:139-150
 *   mocb200_set/get_boundary  <- BoundaryCondition::data_             core/boundary_condition.hpp:144-164
 *   mocb200_sweep             <- MoCSweeper::sweep + sweep1g<CW>      sweepers/moc/moc_sweeper.cpp:189-225,
 *                                + SourceIsotropic::self_scatter      sweepers/moc/moc_sweeper_kernel.inc.hpp:36-180,
 *                                + BoundaryCondition::update          core/source_isotropic.cpp:21-57,
 *                                                                     core/boundary_condition.cpp:145-191
 *   mocb200_get_coarse        <- moc::Current::post_ray tallies       sweepers/moc/moc_current_worker.hpp:202-264
 *   mocb200_set_sweep_inputs,
 *   mocb200_get_sweep_results,
 *   mocb200_pack_results_device <- everything MoCSweeper::sweep reads    sweepers/moc/moc_sweeper.cpp:197-219
 *                                from / leaves in source_, flux_(:, g), boundary_[plane], coarse_data_
 *                                (the single-array calls above, fused: one copy each way, one synchronisation)
 *   mocb200_set_sn_xs,
 *   mocb200_get_corrections   <- cmdo::CurrentCorrections                sweepers/cmdo/correction_worker.hpp:109-246,
 *                                                                     sweepers/cmdo/correction_worker.cpp:32-158
 *
 * All floating point data is FP64, all indices int32 unless stated (the
 * reference: real_t = double, util/global_config.hpp:27-33).
 */
#ifndef MOCC_B200_H
#define MOCC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOCB200_MAX_POLAR 4 /* polar angles that may share one geometry bundle */

/* status codes */
#define MOCB200_OK 0
#define MOCB200_ERR_INVALID 1 /* bad argument / inconsistent problem description */
#define MOCB200_ERR_CUDA 2    /* CUDA runtime error, see mocb200_last_error */
#define MOCB200_ERR_NO_DEVICE 3
#define MOCB200_ERR_STATE 4   /* call sequence error (e.g. sweep before set_xs) */

/* tally modes of mocb200_sweep (what the LAST inner iteration accumulates) */
#define MOCB200_TALLY_NONE 0        /* moc::NoCurrent */
#define MOCB200_TALLY_CURRENT 1     /* moc::Current: coarse currents + surface flux */
#define MOCB200_TALLY_CORRECTIONS 2 /* cmdo::CurrentCorrections (2D3D) */

/* boundary update schemes (MoCSweeper attribute boundary_update, moc_sweeper.cpp:105-116) */
#define MOCB200_BOUNDARY_GS 0     /* reference default; two phases (octant 1+3, then 2+4) */
#define MOCB200_BOUNDARY_JACOBI 1 /* all angles from the previous sweep's outgoing flux */

/* exponential evaluation */
#define MOCB200_EXP_TABLE 0    /* Exponential_Linear<N> table in shared memory, bit-identical entries */
#define MOCB200_EXP_FACTORED 1 /* reserved (not implemented: mocb200_create rejects it) */

/* sweep kernel selection (diagnostics / A-B measurements) */
#define MOCB200_KERNEL_AUTO 0   /* RCHUNK when the attenuation cache fits in device memory, else TRACK */
#define MOCB200_KERNEL_ITEM 1   /* one thread per (track, direction, group), serial walk */
#define MOCB200_KERNEL_TRACK 2  /* one warp per track, both directions, affine scan over lanes */
#define MOCB200_KERNEL_CACHED 3 /* TRACK with the table lookups cached in HBM per cross-section upload */
#define MOCB200_KERNEL_CHUNK 4  /* CACHED with one scan per track: TMA-staged track, lane-owned contiguous chunks */
#define MOCB200_KERNEL_RCHUNK 5 /* CHUNK, second generation: register-resident chunks of fixed length packed into
                                   equal batches, one lane per (chunk, polar angle), one scan per batch */

/*
 * Flattened ray-tracing data ("MOCFLAT"), produced once on the host from the
 * reference's RayData / CoreMesh / AngularQuadrature objects
 * (mocc_b200/host/flatten.cpp). Field names equal the array names of the
 * .mocflat container so that any FFI can fill this struct mechanically.
 *
 * Angle indexing follows the reference: sweep angle a in [0, n_ang) covers
 * octants 1-2 (n_ang = 2*ndir_oct), its reverse is a + n_ang; boundary storage
 * covers n_ang_bc = 4*ndir_oct angles.
 */
typedef struct mocb200_problem {
    /* ---- scalars ---- */
    int32_t n_group;
    int32_t n_reg;        /* FSRs over all macroplanes (MeshTreatment::PLANE) */
    int32_t n_plane;      /* macroplanes */
    int32_t n_unique;     /* geometrically unique planes (ray sets) */
    int32_t ndir_oct;
    int32_t n_ang;        /* 2*ndir_oct */
    int32_t n_geom;       /* distinct ray geometries among the n_ang sweep angles */
    int32_t bc_per_group; /* boundary values per group per plane */
    int32_t n_surf;       /* coarse surfaces, whole mesh */
    int32_t n_cell;       /* coarse cells, whole mesh */
    int32_t n_surf_plane; /* coarse surfaces per plane */
    int32_t n_cell_plane; /* coarse cells per plane (nx*ny) */
    int32_t nx, ny, nz;   /* coarse mesh dimensions */
    int32_t exp_n;        /* Exponential_Linear<N>: intervals (10000) */
    double exp_min;       /* -10.0 */
    double exp_max;       /*   0.0 */
    int64_t n_trk;        /* tracks: unique (unique plane, geometry, ray) */
    int64_t n_seg;        /* segments over all tracks */
    int64_t n_cm;         /* coarse-ray records over all tracks */

    /* ---- per sweep angle [n_ang] ---- */
    const int32_t *ang_geom;     /* geometry id of the angle */
    const double *ang_rsintheta; /* Direction::rsintheta (long double trig), kernel:79 */
    /* ---- per (macroplane, sweep angle) [n_plane*n_ang] ---- */
    const double *wt_v_st;       /* weight*spacing*height*sin(theta)*PI, kernel:80-82 */
    const double *cur_wx, *cur_wy; /* moc::Current current_weights_[0/1], moc_current_worker.hpp:192-195 */
    const double *flx_wx, *flx_wy; /* flux_weights_[0/1], :196-197 */

    /* ---- boundary-condition layout [n_ang_bc = 2*n_ang] ---- */
    const int32_t *bc_offset;   /* start of the angle's block inside one group, boundary_condition.cpp:63-74 */
    const int32_t *bc_size_x;   /* slots on the X-normal face (= RayData::ny) */
    const int32_t *bc_size_y;   /* slots on the Y-normal face (= RayData::nx) */
    const int32_t *bc_dst_off;  /* [n_ang_bc*2] offset of the face this angle's outgoing face (X,Y) feeds */
    const int32_t *bc_dst_kind; /* [n_ang_bc*2] 0 vacuum (write 0), 1 reflect (copy), 2 prescribed (keep) */

    /* ---- geometry: tracks ---- */
    const int64_t *geom_trk_begin; /* [n_unique*n_geom + 1] CSR over tracks */
    const int64_t *trk_seg_begin;  /* [n_trk + 1] CSR over segments */
    const int32_t *trk_bc;         /* [n_trk*2] Ray::bc(0), Ray::bc(1) */
    const int64_t *trk_cm_begin;   /* [n_trk + 1] CSR over coarse-ray records */
    const int32_t *trk_cm_start;   /* [n_trk*4] cm_cell_fw, cm_cell_bw, cm_surf_fw, cm_surf_bw (plane-local) */
    const double *seg_len;         /* [n_seg] Ray::seg_len (after volume correction) */
    const int32_t *seg_fsr;        /* [n_seg] Ray::seg_index (plane-local FSR) */
    const uint32_t *cm_data;       /* [n_cm] fw | bw<<4 | nseg_fw<<8 | nseg_bw<<16 (Ray::RayCoarseData) */

    /* ---- macroplanes [n_plane] ---- */
    const int32_t *plane_unique;      /* ray-set id, MoCSweeper::macroplane_unique_ids_ */
    const int32_t *plane_first_reg;   /* first_reg_macroplane_ */
    const int32_t *plane_cell_offset; /* Mesh::coarse_cell_offset(iplane) */
    const int32_t *plane_surf_offset; /* Mesh::coarse_surf_offset(iplane) */

    /* ---- coarse mesh, one plane [n_cell_plane*4], surfaces E,N,W,S ---- */
    const int32_t *coarse_surf; /* Mesh::coarse_surf(cell, s), plane-local */
    const int32_t *coarse_nbr;  /* Mesh::coarse_neighbor(cell, s), -1 outside */

    /* ---- FSR data [n_reg] ---- */
    const double *vol; /* TransportSweeper::vol_ */

    /* ---- exponential table [exp_n + 2]: exp(exp_min + i*space), last entry repeated ---- */
    const double *exp_table;

    /* ---- 2D3D correction factors (cmdo::CurrentCorrections, correction_worker.cpp:32-158).
     *      Optional: all five NULL disables MOCB200_TALLY_CORRECTIONS. ---- */
    const double *ang_area_x; /* [n_ang] abs(spacing/cos(alpha)), correction_worker.cpp:78-79 */
    const double *ang_area_y; /* [n_ang] abs(spacing/sin(alpha)) */
    const double *ang_ox;     /* [n_ang] Angle::ox (only its sign is used, :47-66) */
    const double *cell_dx;    /* [n_cell_plane] Mesh::pin_dx()[coarse_position(cell).x] */
    const double *cell_dy;    /* [n_cell_plane] Mesh::pin_dy()[coarse_position(cell).y] */
} mocb200_problem;

typedef struct mocb200_sweeper mocb200_sweeper; /* opaque */

/* Options fixed at creation */
typedef struct mocb200_options {
    int32_t device;          /* CUDA device ordinal */
    int32_t boundary_update; /* MOCB200_BOUNDARY_* */
    int32_t exp_mode;        /* MOCB200_EXP_* */
    int32_t max_polar;       /* polar angles bundled per work item (1..MOCB200_MAX_POLAR); 0 = default */
    int32_t block_threads;   /* 0 = default */
    int32_t plane_begin;     /* this rank's macroplane range [plane_begin, plane_end); both 0 = all */
    int32_t plane_end;
    int32_t kernel;          /* MOCB200_KERNEL_* */
    int32_t chunk_cap;       /* CHUNK kernel test hook: cap on the segments staged at once (0 = what fits), forces the
                                super-block chaining of long tracks; negative = that cap with two-warp teams */
    int32_t cache_groups;    /* energy groups the attenuation cache holds at once: 0 = all if they fit in device memory,
                                else as many as fit (>= 1; rebuilt per sweep call, 2 small launches per group); test hook */
    int32_t persistent;      /* RCHUNK kernel: 0 = one launch per boundary phase and inner (default); 1 = all plain inners
                                of a mocb200_sweep call in ONE cooperative launch (grid barriers between the phases,
                                flux / q-bar update inside); 2 = 1 + the first batch behind every barrier requested in
                                front of it. Same results; needs one track list per phase (2-D problems, 3-D problems with
                                one unique plane), otherwise ignored. Environment MOCB200_RC_PERSIST overrides (A/B hook) */
    int32_t family_begin;    /* this rank's ANGLE-FAMILY range [family_begin, family_end) of mocb200_angle_families();
                                both 0 = all. A family is closed under track reversal, polar bundling and the boundary
                                update, so a handle sweeps its families exactly as the whole sweep does; the FSR tally
                                (and the coarse tallies) must be summed over the ranks between mocb200_sweep_partial and
                                mocb200_finalize_flux (moc_sweeper_kernel.inc.hpp:155-173 split at the reduction) */
    int32_t family_end;
    int32_t reserved[3];
} mocb200_options;

/* Build the device-resident problem. The host arrays may be freed afterwards. */
int mocb200_create(const mocb200_problem *prob, const mocb200_options *opt, mocb200_sweeper **out);
int mocb200_destroy(mocb200_sweeper *h);
/* Message for the last non-OK status of this handle (or of create when h is NULL). */
const char *mocb200_last_error(const mocb200_sweeper *h);
/* Use a caller-provided cudaStream_t for all subsequent work (NULL = handle's own stream). */
int mocb200_set_stream(mocb200_sweeper *h, void *cuda_stream);
int mocb200_synchronize(mocb200_sweeper *h);

/*
 * Per-group data. Host arrays are [g_count][n_reg] (one contiguous column per
 * group, groups g_begin .. g_begin+g_count-1).
 *   xstr      transport XS the sweep attenuates with (ExpandedXS incl. TL splitting)
 *   xstr_src  transport XS used in the source normalisation q/(4 pi xstr_src); the
 *             reference uses the UN-split XS there (source_isotropic.cpp:29); NULL = xstr
 *   xs_self   within-group scattering XS Sigma_s(g->g) per FSR (source_isotropic.cpp:30)
 */
int mocb200_set_xs(mocb200_sweeper *h, int g_begin, int g_count, const double *xstr,
                   const double *xstr_src, const double *xs_self);
/* 1-group source WITHOUT self scatter (fission + in-scatter + external), Source::get() */
int mocb200_set_source(mocb200_sweeper *h, int g_begin, int g_count, const double *src);
int mocb200_set_flux(mocb200_sweeper *h, int g_begin, int g_count, const double *flux);
int mocb200_get_flux(mocb200_sweeper *h, int g_begin, int g_count, double *flux);
int mocb200_get_source(mocb200_sweeper *h, int g_begin, int g_count, double *src);
/*
 * SOURCE CONSTRUCTION ON THE DEVICE: what FixedSourceSolver::step does on the host before every
 * sweep(group) (src/solvers/fixed_source_solver.cpp:102-117) and EigenSolver::step before that
 * (src/solvers/eigen_solver.cpp:218-245), from the flux resident on the device -- no source or flux upload per group.
 *   mocb200_set_source_xs        cross sections by cross-section-mesh region: fsr_mat [n_reg] region of every FSR,
 *                                xsnf / xsch [n_mat][n_group], scat [n_mat][to][from], scat_band [n_mat][to][2] = the
 *                                row's ScatteringRow::min_g / max_g (NULL: first / last non-zero entry)
 *   mocb200_set_external_source  Source::add_external's [n_group][n_reg] array (NULL: none; source.cpp:41-54)
 *   mocb200_fission_source       TransportSweeper::calc_fission_source (src/core/transport_sweeper.cpp:119-134):
 *                                fs = sum_g (1/k nu-Sigma_f,g) flux_g from the resident flux
 *   mocb200_set/get_fission_source   the host's array instead / for the host's convergence test
 *   mocb200_build_source         Source::initialize_group + fission + in_scatter (src/core/source.cpp:41-112) for groups
 *                                [g_begin, g_begin + g_count) into the source mocb200_sweep reads. Group g sees the flux
 *                                of the groups swept before it (Gauss-Seidel over groups) when called group by group
 * Same operations in the same order as the reference, not contracted: bit-identical to its host arrays.
 */
int mocb200_set_source_xs(mocb200_sweeper *h, int n_mat, const int32_t *fsr_mat, const double *xsnf, const double *xsch,
                          const double *scat, const int32_t *scat_band);
int mocb200_set_external_source(mocb200_sweeper *h, const double *ext);
int mocb200_fission_source(mocb200_sweeper *h, double k);
int mocb200_set_fission_source(mocb200_sweeper *h, const double *fs);
int mocb200_get_fission_source(mocb200_sweeper *h, double *fs);
int mocb200_build_source(mocb200_sweeper *h, int g_begin, int g_count);
/* Directly impose q-bar (skips self scatter on the next sweep with n_inner == 1 and
 * use_qbar != 0); for sweep1g-level parity tests. [g_count][n_reg] */
int mocb200_set_qbar(mocb200_sweeper *h, int g_begin, int g_count, const double *qbar);

/* Incoming boundary flux of one macroplane: [g_count][bc_per_group], reference layout. */
int mocb200_set_boundary(mocb200_sweeper *h, int plane, int g_begin, int g_count, const double *bc);
int mocb200_get_boundary(mocb200_sweeper *h, int plane, int g_begin, int g_count, double *bc);

/*
 * The hot path. For groups [g_begin, g_begin+g_count) run n_inner inner iterations:
 *   q-bar = (src + flux*xs_self) / (4 pi xstr_src)   (unless use_qbar)
 *   transport sweep over all rays/angles/planes of this handle, boundary update
 *   flux = tally/(xstr*vol) + 4 pi q-bar
 * The last inner also accumulates the tallies selected by tally_mode.
 * Asynchronous on the handle's stream.
 */
int mocb200_sweep(mocb200_sweeper *h, int g_begin, int g_count, int n_inner, int tally_mode, int use_qbar);

/* The transfers of one sweep(group) call fused (one staging copy each way, ONE synchronisation):
 * what MoCSweeper::sweep reads from / leaves in the host objects around sweep1g
 * (moc_sweeper.cpp:197-219: source_->get(), flux_(:, group), boundary_[plane]).
 *   set_sweep_inputs : source[n_reg], flux[n_reg] (either may be NULL = keep the device copy) and, for every
 *                      macroplane ip of the mesh, boundary[ip] -> [bc_per_group] incoming boundary flux
 *                      (boundary NULL or entry NULL = keep); planes outside this handle's range are ignored.
 *                      Returns once the host arrays may be reused; the copy is stream-ordered before the sweep.
 *   get_sweep_results: flux[n_reg] (only this handle's FSR range is written), boundary[ip] as above,
 *                      current/surface_flux[n_surf] as mocb200_get_coarse (both NULL = skip). Synchronises.
 * Page-locked host arrays (cudaHostAlloc / cudaHostRegister) are copied to / from directly, without the staging
 * memcpy; such input arrays must stay unmodified until the next synchronising call on the handle. */
int mocb200_set_sweep_inputs(mocb200_sweeper *h, int group, const double *source, const double *flux,
                             const double *const *boundary);
int mocb200_get_sweep_results(mocb200_sweeper *h, int group, double *flux, double *const *boundary, double *current,
                              double *surface_flux);

/*
 * Multi-rank exchange (one process per GPU, planes sharded by plane_begin/plane_end): packs this handle's part of
 * the results of `group` into a DEVICE buffer of the caller, on the handle's stream, with no host round trip:
 * [scalar flux of its FSR range][radial coarse current of its planes' surfaces][their surface flux]. The buffer
 * is what the ranks exchange with one ncclAllGather per sweep(group) -- the only inter-GPU traffic of the path
 * (replaces, across processes, what CoarseData / flux_ hold in the single-process reference: coarse_data.hpp:151-152,
 * transport_sweeper.hpp:139-150). dst_device == NULL: *count receives the doubles the handle would write.
 */
int mocb200_pack_results_device(mocb200_sweeper *h, int group, double *dst_device, size_t capacity, size_t *count);

/* Raw radial coarse tallies of the last TALLY_CURRENT/CORRECTIONS sweep for one group:
 * current[n_surf], surface_flux[n_surf] (whole-mesh surface indexing, x/y-normal surfaces of
 * this handle's macroplanes only, NOT yet divided by the surface area -- the reference does
 * that in post_sweep, moc_current_worker.hpp:272-318). */
int mocb200_get_coarse(mocb200_sweeper *h, int group, double *current, double *surface_flux);

/*
 * 2D3D coupling (MoCSweeper_2D3D, cmdo/moc_sweeper_2d3d.cpp:48-99).
 *   mocb200_set_sn_xs        homogenised transport XS of the Sn mesh the beta factor divides by
 *                            (xstr_sn_[cell + mplane_offset], correction_worker.cpp:84-93): host array
 *                            [g_count][n_plane*n_cell_plane], macroplane-major.
 *   mocb200_get_corrections  correction factors of the last MOCB200_TALLY_CORRECTIONS sweep of `group`:
 *                            alpha[2*n_ang][n_plane*n_cell_plane][2] (X, Y normal), beta[2*n_ang][n_plane*n_cell_plane],
 *                            angle index = sweep angle (forward) or sweep angle + n_ang (its reverse), cell index =
 *                            cell + Mesh::coarse_cell_offset(macroplane). Entries of macroplanes outside this handle's
 *                            range are left untouched.
 */
int mocb200_set_sn_xs(mocb200_sweeper *h, int g_begin, int g_count, const double *xs);
int mocb200_get_corrections(mocb200_sweeper *h, int group, double *alpha, double *beta);

/* Counters for bench/diagnostics */
typedef struct mocb200_stats {
    int64_t kernel_launches;   /* kernels launched by this handle since creation */
    int64_t sweep_launches;    /* of which transport-sweep kernels */
    int64_t segments_per_sweep; /* reference segment count S (per-polar copies counted), this handle's planes */
    int64_t unique_segments;   /* segments resident on the device */
    int64_t device_bytes;      /* device memory held */
    int64_t items[2];          /* work items (track, direction) per boundary phase */
    int64_t kernel;            /* MOCB200_KERNEL_* in use (AUTO resolved) */
    int64_t swept_segments;    /* segments the track lists hold per plane set (polar copies bundled) */
} mocb200_stats;
int mocb200_get_stats(const mocb200_sweeper *h, mocb200_stats *out);

/* Time (ms, CUDA events on the handle's stream) of the sweep kernels of the LAST inner iteration of the last
 * mocb200_sweep call (the tallying one when a tally was requested; with mocb200_options.persistent and no tally: the
 * whole persistent launch); synchronises. For the time of every inner use mocb200_set_timing / mocb200_get_timing. */
int mocb200_last_sweep_ms(mocb200_sweeper *h, double *ms);

/*
 * ANGLE-FAMILY SHARDING of one plane over ranks (single-plane 2-D cases). Replaces nothing in the
 * reference (its sweep is one shared-memory loop over the angles, moc_sweeper_kernel.inc.hpp:59); what it splits is
 * that loop: every rank sweeps the angles of its families, the per-FSR tally t_flux is summed over the ranks where the
 * reference sums it over its threads (:155-163), then the flux update (:165-173) runs on every rank.
 *   mocb200_angle_families   families of a problem (no device needed); family_of_angle: [2 n_ang] or NULL
 *   mocb200_sweep_partial    ONE inner sweep (self scatter, sweep of the handle's angles, coarse tallies of those angles
 *                            when tally_mode says so) that leaves the tally un-normalised
 *   mocb200_device_buffer    device address and length (doubles) of what has to be summed over the ranks:
 *                            MOCB200_BUF_TALLY (group-major [group - g_begin][stride], the swept groups first),
 *                            MOCB200_BUF_CURRENT / _SURFACE_FLUX ([n_surf][8-padded groups])
 *   mocb200_adopt_device_buffer  make the handle use caller-owned device memory for one of them (e.g. a tensor of the
 *                            communication library: all-reduce in place, no copies); contents are carried over
 *   mocb200_finalize_flux    flux = tally / (xstr vol) + 4 pi q-bar for the groups of the partial sweep
 */
#define MOCB200_BUF_TALLY 0
#define MOCB200_BUF_CURRENT 1
#define MOCB200_BUF_SURFACE_FLUX 2
int mocb200_angle_families(const mocb200_problem *prob, int32_t *n_family, int32_t *family_of_angle);
int mocb200_sweep_partial(mocb200_sweeper *h, int g_begin, int g_count, int tally_mode, int use_qbar);
int mocb200_finalize_flux(mocb200_sweeper *h, int g_begin, int g_count);
int mocb200_device_buffer(mocb200_sweeper *h, int which, void **ptr, int64_t *count);
int mocb200_adopt_device_buffer(mocb200_sweeper *h, int which, void *ptr, int64_t count);

/* Cumulative device time of the transport-sweep kernels. While enabled, every inner iteration
 * of mocb200_sweep is bracketed by a CUDA event pair on the handle's stream;
 * mocb200_get_timing synchronises, returns the summed time (ms) and the number of inner sweeps
 * since the last call, and resets both. */
int mocb200_set_timing(mocb200_sweeper *h, int enabled);
int mocb200_get_timing(mocb200_sweeper *h, double *sweep_ms, int64_t *inner_sweeps);

/* Library/version string */
const char *mocb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MOCC_B200_H */
