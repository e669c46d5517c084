/*
 * c2ray_oracle.h -- CPU restatement of the C2-Ray3Dm photo-ionization hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (c2ray3dm_b200/, the C-ABI library)
 * may include, link or call this.  Allowed users: tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or expected outputs, and it
 * cannot be compiled here (no Fortran compiler in the image), so this restatement is pinned
 * only by review against the cited reference lines plus the invariants in tests/ (see DESIGN.md).
 *
 * All "file:line" citations are relative to the reference tree (garrelt/C2-Ray3Dm).
 */
#ifndef C2RAY_ORACLE_H
#define C2RAY_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NUMTAU 2000   /* radiation_sizes.f90:14 */
#define ORC_NUMFREQ 128   /* radiation_sizes.f90:13 */

/* The values of the reference's Fortran parameters as the compiler sees them
 * (single-precision literals widened to double). */
typedef struct orc_constants {
  double pi, sigma_HI_at_ion_freq, eth0, ev2k, temph0, colh0, ev2fr, ion_freq_HI, ion_freq_HeII;
  double bb_MaxFreq, two_pi_over_c_square, bh00, albpow, hplanck, k_B, c_light, m_p, sigma_SB;
  double abu_he, abu_c, mu, h, Omega0, Omega_B, Mpc, H0, rho_crit_0, YEAR, R_SOLAR;
  double epsilon, convergence_fraction, minimum_fractional_change, minimum_fraction_of_atoms;
  double loss_fraction, max_coldensh, tau_photo_limit, sqrt3, sqrt2, minlogtau, dlogtau;
  double xh_initial, bb_Teff, bb_S_star;
} orc_constants;

void orc_get_constants(orc_constants *c);

/* rad_ini (radiation_tables.F90:95-126): fills thick/thin tables [0..NumTau] and diagnostics. */
typedef struct orc_rad_diag {
  double S_star_unscaled, S_scaling, R_star, h_over_kT, freq_min, freq_max, delta_freq;
  double romw7[ORC_NUMFREQ + 1];
} orc_rad_diag;
void orc_rad_ini(double *thick, double *thin, orc_rad_diag *diag);

/* Module-global state of the reference that the hot path reads and writes. */
typedef struct orc_state orc_state;

orc_state *orc_create(int m1, int m2, int m3);
void orc_destroy(orc_state *s);

void orc_set_tables(orc_state *s, const double *thick, const double *thin);
void orc_set_density(orc_state *s, const float *ndens);
void orc_set_geometry(orc_state *s, const double dr[3], double vol);
void orc_set_clumping(orc_state *s, int type_of_clumping, float clumping, const float *grid);
void orc_deterministic_clumping(orc_state *s, double p1, double p2, double p3, double avg_dens);
const float *orc_clumping_grid(const orc_state *s);
void orc_set_lls(orc_state *s, int use_LLS, int type_of_LLS, double coldensh_LLS,
                 const float *grid, double R_max_LLS);
void orc_set_temperature(orc_state *s, double temper_val);
void orc_set_sources(orc_state *s, int NumSrc, const int32_t *srcpos /*3 x NumSrc, 1-based*/,
                     const double *NormFlux_stellar /*NumSrc*/, double S_star);
void orc_set_xh(orc_state *s, const double *xh);
void orc_set_xh_av(orc_state *s, const double *x);
void orc_set_loss_fraction(orc_state *s, double loss_fraction);
/* walk order for do_source: 0 = serial evolve2D (nthreads==1 branch), 1 = axes/planes/octants
 * (the OpenMP branch executed serially), 2 = Chebyshev shells with a scrambled in-shell order. */
void orc_set_walk_order(orc_state *s, int order);
/* MPI emulation: this "rank" does sources 1+rank, 1+rank+npr, ... (master_slave.F90:85) */
void orc_set_rank(orc_state *s, int rank, int npr);
/* number of OpenMP threads used by orc_pass_all_sources in source-parallel CPU-baseline mode. */
void orc_set_threads(orc_state *s, int nthreads);
/* 0 (default): one source per thread, private rate grids summed (do_grid_static + MPI_ALLREDUCE);
 * 1: all threads inside one source as in the OpenMP build (evolve_source.F90:141-186) */
void orc_set_omp_in_source(orc_state *s, int on);

/* ---- non-isothermal path (isothermal=.false.; thermal.f90, cooling.f90, heat_lookuptable) ---------- */
void orc_rad_ini_heat(double *thick, double *thin, double *heat_thick, double *heat_thin);
void orc_set_isothermal(orc_state *s, int isothermal);   /* after orc_set_temperature: grids start at temper_val */
void orc_set_heat_tables(orc_state *s, const double *heat_thick, const double *heat_thin);
void orc_set_cooling_table(orc_state *s, const double *log10_temp /*61*/, const double *log10_cool /*61*/);
void orc_set_redshift(orc_state *s, double zred, int cosmological);   /* cosmology.F90:42, c2ray_parameters.f90:105 */
float *orc_temperature_grid(orc_state *s);   /* 3 x ncell: (current, average, intermed) per cell */
double *orc_phiheat(orc_state *s);

double *orc_xh(orc_state *s);
double *orc_xh_av(orc_state *s);
double *orc_xh_intermed(orc_state *s);
double *orc_phih(orc_state *s);
double *orc_coldensh_out(orc_state *s);

/* column_density.f90:29-271 on the state's coldensh_out; pos/srcpos are 1-based (pos unwrapped). */
void orc_cinterp(const orc_state *s, const int pos[3], const int srcpos[3], double *cdensi,
                 double *path);

/* radiation_photoionrates.F90:71-317 for one stellar source: returns photo_cell_HI, photo_in,
 * photo_out through out[3]. */
void orc_photoion_rates(const orc_state *s, double colum_in, double colum_out, double vol,
                        double normflux, double out[3]);

/* doric.f90:33-134 */
void orc_doric(const orc_state *s, double dt, double temp0, double rhe, double rhh, double xfh[2],
               double xfh_av[2], double phih, float clumping);

typedef struct orc_source_report {
  int nbox;
  double photon_loss_src;
  int64_t updates; /* evolve0D calls that passed the coldensh_out==0 gate */
} orc_source_report;

/* evolve_source.F90:58-221 for source ns (1-based). Adds into phih_grid. */
void orc_do_source(orc_state *s, int ns, orc_source_report *rep);

typedef struct orc_pass_report {
  double photon_loss_all;
  int64_t sum_nbox_all;
  int64_t updates;
} orc_pass_report;

/* evolve.F90:430-440 + 444-495 (set_rates_to_zero is NOT included; call orc_set_rates_to_zero). */
void orc_set_rates_to_zero(orc_state *s);
void orc_pass_all_sources(orc_state *s, orc_pass_report *rep);

typedef struct orc_photon_stats {
  double h0_before, h1_before, h0_after, h1_after;
  double totrec, totcollisions, dh0, total_ion;
  double totalsrc, photcons, total_photon_loss, LLS_loss;
} orc_photon_stats;

typedef struct orc_global_report {
  int conv_flag;
  double min_avg_neutral; /* 1.0-maxval(xh_av) before the pass, evolve.F90:535 */
  double sum_xh_intermed; /* after the pass */
  orc_photon_stats stats;
} orc_global_report;

/* evolve.F90:499-573 */
void orc_global_pass(orc_state *s, double dt, double photon_loss_all, orc_global_report *rep);

void orc_state_before(orc_state *s);                                      /* photonstatistics.F90:104-132 */
void orc_calculate_photon_statistics(orc_state *s, double dt, const double *xh_l,
                                     const double *xh_r, orc_photon_stats *st); /* :82-99 */

#define ORC_MAX_ITER 104
typedef struct orc_step_report {
  int niter;
  int converged; /* 1 = "Multiple sources convergence reached", 0 = "not converging" */
  int conv_criterion;
  int conv_flag[ORC_MAX_ITER];
  double rel_change_sum_xh1[ORC_MAX_ITER], rel_change_sum_xh0[ORC_MAX_ITER];
  double photon_loss_all[ORC_MAX_ITER];
  int64_t sum_nbox_all[ORC_MAX_ITER];
  int64_t updates[ORC_MAX_ITER];
  orc_photon_stats iter_stats[ORC_MAX_ITER];
  orc_photon_stats final_stats;
  double grtotal_ion, grtotal_src;
  int64_t total_updates;
  double seconds_raytrace, seconds_global;
} orc_step_report;

/* evolve.F90:83-281 with restart==0.  max_outer_iter<=0 means the reference's limit (niter>100). */
void orc_evolve3D(orc_state *s, double dt, int max_outer_iter, orc_step_report *rep);

/* write_iteration_dump / start_from_dump, evolve.F90:285-426.  The reference dumps between pass_all_sources and
 * global_pass once 15 minutes have passed (:253-266); here the caller names the iteration (0 = never) and reads
 * the record (niter | photon_loss_all | phih_grid | xh_av | xh_intermed, :309-317) back with orc_get_dump.
 * orc_evolve3D_restart is evolve3D(time,dt,restart/=0): the record replaces iterdump[12].bin, xh must hold the
 * state of the start of the step, and global_pass runs once before the loop continues at niter+1 (:154-158). */
void orc_set_dump_iteration(orc_state *s, int niter);
int orc_get_dump(const orc_state *s, int *niter, double *photon_loss_all, double *phih, double *xh_av,
                 double *xh_intermed);
void orc_evolve3D_restart(orc_state *s, double dt, int max_outer_iter, int niter, double photon_loss_all,
                          const double *phih, const double *xh_av, const double *xh_intermed,
                          orc_step_report *rep);

#ifdef __cplusplus
}
#endif
#endif
