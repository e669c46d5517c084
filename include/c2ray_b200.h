/*
 * c2ray_b200.h -- C ABI of the B200-native C2-Ray3Dm photo-ionization hot path.
 *
 * This library replaces the body of `module evolve` (reference evolve.F90:61,76,83 --
 * `subroutine evolve3D(time,dt,restart)`, called from exactly one site, C2Ray.F90:379) and
 * everything below it: pass_all_sources/do_grid/do_source/evolve0D/cinterp/photoion_rates
 * (evolve.F90:444-495, master_slave.F90:53-96, evolve_source.F90:58-221, evolve_point.F90:83-299,
 * column_density.f90:29-293, radiation_photoionrates.F90:71-317), global_pass/evolve0D_global/
 * do_chemistry/doric (evolve.F90:499-573, evolve_point.F90:305-555, doric.f90:33-134), the photon
 * statistics (photonstatistics.F90:82-293) and the rank reduction (evolve.F90:577-616).
 *
 * The reference passes its inputs as Fortran module state; the setters below marshal exactly
 * that state.  Conventions:
 *   - plain C linkage, POD arguments, no C++/torch types;
 *   - arrays are Fortran column-major (i fastest) flat host pointers unless the name ends in
 *     `_dev`; the caller owns every host buffer, the library copies and never retains it;
 *   - source positions are 1-based mesh indices exactly as in srcpos(3,NumSrc) (sourceprops.F90:56);
 *   - every function returns 0 on success, non-zero on failure; c2b_last_error() has the text;
 *     nothing aborts or throws across this boundary;
 *   - a handle is bound to one CUDA device and is not thread-safe; one handle per process/GPU.
 *     With nranks > 1 each handle traces sources ns = 1+rank, 1+rank+nranks, ... (the split of
 *     do_grid_static, master_slave.F90:85) and the partial rate grids are summed with NCCL.
 *   - there is no CPU fallback: without a CUDA device c2b_create() fails.
 */
#ifndef C2RAY_B200_H
#define C2RAY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C2B_NUMTAU 2000      /* radiation_sizes.f90:14 */
#define C2B_MAX_ITER 104     /* evolve.F90:228: outer iterations stop after niter > 100 */
#define C2B_UNIQUE_ID_BYTES 128

typedef struct c2b_handle c2b_handle;

/* Compile-time parameters of the reference that reach the hot path (SURVEY 8b "hidden inputs").
 * c2b_default_config() fills in the values of the reference as shipped, with the Fortran
 * literal semantics (default-real literals widened to double). */
typedef struct c2b_config {
  int32_t mesh[3];            /* sizes.f90:33 */
  int32_t device;             /* CUDA device ordinal */
  int32_t rank, nranks;       /* my_mpi: rank, npr */
  int32_t isothermal;         /* c2ray_parameters.f90:28; 0 = heating and cooling (thermal.f90), needs the heat
                                 tables, a cooling table, the redshift and a temperature grid (below) */
  int32_t type_of_clumping;   /* :75   1,2 = scalar; 3,4,5 = float32 grid */
  int32_t use_LLS;            /* :80 */
  int32_t type_of_LLS;        /* :87   1 = scalar, 2 = float32 grid, 3 = R_max barrier */
  int32_t subboxsize;         /* :54 */
  int32_t max_subbox;         /* :61 */
  int32_t max_outer_iter;     /* evolve.F90:228 (100) */
  int32_t reserved0;
  double epsilon;                    /* c2ray_parameters.f90:31 */
  double convergence_fraction;       /* :25 */
  double minimum_fractional_change;  /* :34 */
  double minimum_fraction_of_atoms;  /* :40 */
  double loss_fraction;              /* :67 */
  double max_coldensh;               /* evolve_point.F90:95 */
  double tau_photo_limit;            /* radiation_photoionrates.F90:244 */
  double minlogtau, dlogtau;         /* radiation_tables.F90:45-47 */
  double sigma_HI;                   /* radiation_sizes.f90:77 = cgsphotoconstants.f90:24 */
  double pi;                         /* mathconstants.f90:21 */
  double sqrt2, sqrt3;               /* column_density.f90:52-53 */
  double bh00, albpow, colh0, temph0;/* cgsconstants.f90:64-86 */
  double abu_c;                      /* abundances.f90:26 */
  /* non-isothermal path (isothermal == 0): thermal.f90, cooling.f90, heat_lookuptable, cosmo_cool */
  double k_B;                        /* cgsconstants.f90:34 */
  double gamma1;                     /* atomic.f90:23-25: 5/3 - 1 */
  double minitemp;                   /* c2ray_parameters.f90:108 */
  double relative_denergy;           /* :110 */
  double tau_heat_limit;             /* radiation_photoionrates.F90:333 */
  double H0, Omega0;                 /* cosmoparms.f90:30,41 (cosmo_cool, cosmology.F90:198-225) */
  int32_t cosmological;              /* c2ray_parameters.f90:105 */
  int32_t reserved1;
} c2b_config;

/* photonstatistics.F90:41-55 module variables + the report line of :254-281 */
typedef struct c2b_photon_stats {
  double h0_before, h1_before, h0_after, h1_after;
  double totrec, totcollisions, dh0, total_ion;
  double totalsrc, photcons, total_photon_loss, LLS_loss;
} c2b_photon_stats;

/* what pass_all_sources (evolve.F90:444-495) leaves behind */
typedef struct c2b_pass_report {
  double photon_loss_all;   /* evolve.F90:587 / :485 */
  int64_t sum_nbox_all;     /* :612 */
  int64_t updates;          /* evolve0D calls that passed the gate evolve_point.F90:128, all ranks */
  double ms_raytrace;       /* device time of the ray-trace kernels of this rank */
  double ms_allreduce;      /* device time of the rank reduction */
} c2b_pass_report;

/* what global_pass (evolve.F90:499-573) logs and returns */
typedef struct c2b_global_report {
  int32_t conv_flag;        /* number of non-converged points, :558 */
  int32_t reserved0;
  double min_avg_neutral;   /* 1.0-maxval(xh_av) before the pass, :535 */
  double sum_xh_intermed;   /* after the pass, :565 / :183 */
  c2b_photon_stats stats;   /* calculate_photon_statistics(dt,xh_intermed,xh_av), :570 */
  double ms_chemistry;
} c2b_global_report;

/* everything evolve3D logs for one call */
typedef struct c2b_step_report {
  int32_t niter;
  int32_t converged;        /* 1: "Multiple sources convergence reached"; 0: "not converging" */
  int32_t conv_criterion;   /* evolve.F90:162 */
  int32_t reserved0;
  int32_t conv_flag[C2B_MAX_ITER];            /* index = niter (1-based; [0] unused) */
  double rel_change_sum_xh1[C2B_MAX_ITER];    /* index = niter at the test, 0-based */
  double rel_change_sum_xh0[C2B_MAX_ITER];
  double photon_loss_all[C2B_MAX_ITER];
  int64_t sum_nbox_all[C2B_MAX_ITER];
  int64_t updates[C2B_MAX_ITER];
  c2b_photon_stats iter_stats[C2B_MAX_ITER];
  c2b_photon_stats final_stats;               /* evolve.F90:277 */
  double grtotal_ion, grtotal_src;            /* photonstatistics.F90:286-293 */
  int64_t total_updates;
  int64_t kernel_launches;                    /* kernels of this library launched during the call */
  double ms_raytrace, ms_allreduce, ms_chemistry, ms_total; /* device times (CUDA events) */
} c2b_step_report;

/* ---- life cycle ------------------------------------------------------------------------- */
int c2b_default_config(c2b_config *cfg);
int c2b_create(const c2b_config *cfg, c2b_handle **out);  /* replaces evolve_ini, evolve_data.F90:73-93 */
void c2b_destroy(c2b_handle *h);
const char *c2b_last_error(const c2b_handle *h);          /* h may be NULL: last create error */
int c2b_device_count(void);

/* ---- multi-GPU (replaces mpi_setup's communicator for the collectives of evolve.F90:577-616) */
int c2b_get_unique_id(void *id /* C2B_UNIQUE_ID_BYTES */);  /* rank 0; broadcast by the host (MPI_BCAST) */
int c2b_comm_init(c2b_handle *h, const void *id);           /* collective over all nranks handles */

/* ---- inputs: the module state evolve3D reads --------------------------------------------- */
/* rad_ini result: stellar_photo_thick_table(0:NumTau,1), ..thin.. (radiation_tables.F90:78-79) */
int c2b_set_tables(c2b_handle *h, const double *thick, const double *thin, int32_t n /* NumTau+1 */);
/* rad_ini itself for the black-body SED (radiation_tables.F90:95-126), computed on the device */
int c2b_rad_ini_blackbody(c2b_handle *h, double T_eff, double S_star, double freq_min,
                          double freq_max, double hplanck, double k_B,
                          double two_pi_over_c_square, double R_solar,
                          double pl_index_cross_section, double *thick_out, double *thin_out);
int c2b_set_density(c2b_handle *h, const float *ndens);                 /* density_module.F90:22 */
int c2b_set_geometry(c2b_handle *h, const double dr[3], double vol);    /* grid.F90:25,29 */
int c2b_cosmo_evol(c2b_handle *h, double zfactor);                      /* cosmology.F90:161-193 on the device copy */
int c2b_set_clumping_scalar(c2b_handle *h, float clumping);             /* clumping_module.F90:17 */
int c2b_set_clumping_grid(c2b_handle *h, const float *clumping_grid);   /* :18 */
/* deterministic_clumping (clumping_module.F90:327-363, type_of_clumping 3) on the device-resident density:
 * clumping_grid = real(p1*d*d + p2*d + p3) with d = ndens/avg_dens; p1..p3 are the redshift-weighted rows of
 * params_dcm (:350), which the host keeps interpolating (weight_function, :271-307).  Saves the upload of the grid
 * with every slice.  The stochastic model (type 4, :366-438) draws from random_number after random_seed() per cell
 * and is not reproducible by construction: the host keeps generating it and passes it with c2b_set_clumping_grid. */
int c2b_set_clumping_from_density(c2b_handle *h, double p1, double p2, double p3, double avg_dens);
int c2b_get_clumping_grid(c2b_handle *h, float *clumping_grid);
int c2b_set_lls_scalar(c2b_handle *h, double coldensh_LLS);             /* LLS.F90:79 */
int c2b_set_lls_grid(c2b_handle *h, const float *LLS_grid);             /* LLS.F90:81 */
int c2b_set_lls_rmax(c2b_handle *h, double R_max_LLS);                  /* LLS.F90:107 */
int c2b_set_temperature(c2b_handle *h, double temper_val);              /* temperature_module.F90:33 */
int c2b_set_sources(c2b_handle *h, int32_t NumSrc, const int32_t *srcpos /* 3 x NumSrc, 1-based */,
                    const double *NormFlux_stellar /* NumSrc; element 0 = source 1 */,
                    double S_star);                                     /* sourceprops.F90:56-63 */
int c2b_set_xh(c2b_handle *h, const double *xh);                        /* ionfractions_module.F90:22 */
/* ---- non-isothermal inputs (isothermal == 0) ----------------------------------------------- */
/* stellar_heat_thick_table(0:NumTau,1), ..thin.. (radiation_tables.F90:82-83); c2b_rad_ini_blackbody builds them
 * itself when the handle is not isothermal, c2b_get_heat_tables reads them back */
int c2b_set_heat_tables(c2b_handle *h, const double *heat_thick, const double *heat_thin, int32_t n);
int c2b_get_heat_tables(c2b_handle *h, double *heat_thick, double *heat_thin);
/* the 61 rows (log10 T, log10 Lambda) of tables/corocool.tab as setup_cool reads them (cooling.f90:64-87) */
int c2b_set_cooling_table(c2b_handle *h, const double *log10_temp, const double *log10_cool, int32_t n /* 61 */);
int c2b_set_redshift(c2b_handle *h, double zred);                       /* cosmology.F90:42 (cosmo_cool) */
/* temperature_grid(mesh)%(current,average,intermed), default real (temperature_module.F90:21-35): 3 floats per
 * cell, cell index i fastest.  c2b_set_temperature(temper_val) fills all three with temper_val
 * (temperature_array_init, :44-67). */
int c2b_set_temperature_grid(c2b_handle *h, const float *temperature_grid);
int c2b_get_temperature_grid(c2b_handle *h, float *temperature_grid);

/* ---- the hot path ------------------------------------------------------------------------ */
/* coarse: evolve3D(time,dt,restart) evolve.F90:83-281; restart must be 0 unless
 * c2b_set_iter_state() was called (start_from_dump, :328-426) */
int c2b_evolve3d(c2b_handle *h, double time, double dt, int32_t restart, c2b_step_report *rep);

/* fine: lets the Fortran shim keep the outer loop and its log lines (evolve.F90:136-279) */
int c2b_begin_step(c2b_handle *h, double *sum_xh /* sum(xh), for evolve.F90:183 */);  /* :136-153 */
int c2b_pass_all_sources(c2b_handle *h, int32_t niter, double dt, c2b_pass_report *rep); /* :240-246 */
int c2b_global_pass(c2b_handle *h, double dt, c2b_global_report *rep);                   /* :269 */
int c2b_end_step(c2b_handle *h, double dt, int32_t converged, c2b_photon_stats *final_stats); /* :215-217,277-279 */

/* ---- outputs ----------------------------------------------------------------------------- */
int c2b_get_xh(c2b_handle *h, double *xh);
int c2b_get_xh_av(c2b_handle *h, double *xh_av);
int c2b_get_xh_intermed(c2b_handle *h, double *xh_intermed);
int c2b_get_phih(c2b_handle *h, double *phih_grid);
int c2b_get_phih_f32(c2b_handle *h, float *phih_grid_si);  /* real(phih_grid,si), output.F90:359 */
int c2b_get_phiheat(c2b_handle *h, double *phiheat_grid);   /* evolve_data.F90:42 (non-isothermal) */
int c2b_get_source_nbox(c2b_handle *h, int32_t *nbox /* NumSrc; with nranks > 1 every rank holds the counts of all sources */);
int c2b_get_source_loss(c2b_handle *h, double *loss /* NumSrc */);
/* iteration dump / restart (evolve.F90:285-426): niter, photon_loss_all, phih, xh_av, xh_intermed */
int c2b_get_iter_state(c2b_handle *h, int32_t *niter, double *photon_loss_all, double *phih_grid,
                       double *xh_av, double *xh_intermed);
int c2b_set_iter_state(c2b_handle *h, int32_t niter, double photon_loss_all, const double *phih_grid,
                       const double *xh_av, const double *xh_intermed);

/* the two extra records of a non-isothermal dump: phiheat_grid | temperature_grid (evolve.F90:314-317) */
int c2b_get_iter_state_thermal(c2b_handle *h, double *phiheat_grid, float *temperature_grid);
int c2b_set_iter_state_thermal(c2b_handle *h, const double *phiheat_grid, const float *temperature_grid);

/* ---- device-side access for harnesses that already hold data in HBM (bench, tests) -------- */
void *c2b_dev_ptr(c2b_handle *h, const char *name); /* "ndens","xh","xh_av","xh_intermed","phih" */
int c2b_synchronize(c2b_handle *h);
/* keep / bring back a device-resident copy of xh (benchmarks start every step from the same state) */
int c2b_save_xh_dev(c2b_handle *h);
int c2b_restore_xh_dev(c2b_handle *h);
/* single-source diagnostic: traces source ns (1-based) alone into a zeroed phih and returns the
 * full coldensh_out grid the reference would hold after do_source (evolve_source.F90:58-221). */
int c2b_trace_source_debug(c2b_handle *h, int32_t ns, double *coldensh_out, double *phih_grid,
                           int32_t *nbox, double *photon_loss_src);
/* measured peak rate of a dependent-free DFMA stream on this device, in FP64 instructions/s
 * (roofline denominator for the FP64-issue bound) */
int c2b_measure_dfma_rate(c2b_handle *h, double *dfma_per_s);
/* how this handle's sources were dealt to the three ray-trace work-group shapes since c2b_create (diagnostic;
 * the reference has one shape, do_source on one thread, evolve_source.F90:58): counts[0] one CTA per source,
 * [1] one cluster of 6 CTAs per source, [2] one warp per source for the first subbox, [3] of those, handed over to
 * the one-CTA shape because their loss still exceeded loss_fraction (evolve_source.F90:128-131) */
int c2b_get_route_counts(c2b_handle *h, int64_t counts[4]);
/* the rule that deals sources to ranks between passes when nranks > 1 (host logic, no device needed): cost[s] =
 * predicted updates of source s, speed[r] = relative speed of rank r (null: all equal); longest trace first, each to
 * the rank furthest below its share.  The first pass of a source list uses the reference's static rule
 * ns = 1+rank, 1+rank+npr, ... (master_slave.F90:85); afterwards the library calls this with the all-reduced subbox
 * counts and the ranks' measured ray-trace speeds (the reference's dynamic counterpart is do_grid_master,
 * master_slave.F90:124-231). */
int c2b_deal_sources(int32_t nsrc, const int64_t *cost, int32_t nranks, const double *speed, int32_t *owner);
/* owner[s] = rank that traces source s in the next pass (NumSrc entries) */
int c2b_get_source_owner(c2b_handle *h, int32_t *owner);

#ifdef __cplusplus
}
#endif
#endif
