/*
 * c2ray_oracle.c -- CPU restatement of the C2-Ray3Dm photo-ionization hot path (evolve3D).
 *
 * TEST INFRASTRUCTURE ONLY (see c2ray_oracle.h).  PARITY UNPINNED by the reference itself: it
 * has no tests/golden vectors and cannot be compiled in this image (no Fortran compiler).
 *
 * The code follows the reference statement by statement, including the Fortran literal
 * semantics (default-real literals are binary32 and are widened afterwards), evaluation order
 * (left to right, parentheses honoured), and the quirks listed in SURVEY Appendix E.
 * Compile with: gcc -O2 -ffp-contract=off -fno-fast-math [-fopenmp]
 *
 * Citations "file:line" are relative to the reference tree.
 */
#include "c2ray_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* Constants: value = what gfortran stores for the parameter.              */
/* ------------------------------------------------------------------------------------------ */
#define F32(x) ((double)(float)(x)) /* a default-real literal / expression widened to real(dp) */

static const double K_pi = F32(3.141592654f);                 /* mathconstants.f90:21 */
static const double K_m_p = 1.672661e-24;                     /* cgsconstants.f90:26 */
static const double K_c = 2.997925e+10;                       /* :28 */
static const double K_hplanck = 6.6260755e-27;                /* :30 */
static const double K_sigma_SB = 5.670e-5;                    /* :32 */
static const double K_k_B = 1.381e-16;                        /* :34 */
static const double K_albpow = -0.7;                          /* :64 */
static const double K_bh00 = 2.59e-13;                        /* :66 */
static const double K_eth0 = F32(13.598f);                    /* :76 */
static const double K_ethe1 = F32(54.416f);                   /* :101 */
static const double K_abu_he = F32(0.074f);                   /* abundances.f90:23 */
static const double K_abu_c = F32(7.1e-7f);                   /* abundances.f90:26 */
static const double K_epsilon = 1e-14;                        /* c2ray_parameters.f90:31 */
static const double K_convergence_fraction = F32(1.0e-4f);    /* :25 */
static const double K_minimum_fractional_change = F32(1.0e-3f); /* :34 */
static const double K_minimum_fraction_of_atoms = F32(1.0e-8f); /* :40 */
static const int K_subboxsize = 5;                            /* :54 */
static const int K_max_subbox = 1000;                         /* :61 */
static const double K_max_coldensh = F32(2e19f);              /* evolve_point.F90:95 */
static const double K_tau_photo_limit = F32(1.0e-7f);         /* radiation_photoionrates.F90:244 */
static const double K_tau_heat_limit = F32(1.0e-4f);         /* radiation_photoionrates.F90:333 */
static const double K_minitemp = F32(1.0f);                   /* c2ray_parameters.f90:108 */
static const double K_relative_denergy = F32(0.1f);           /* :110 */
static const double K_gamma1 = 5.0 / 3.0 - 1.0;               /* atomic.f90:23-25 */
static const double K_Omega0 = F32(0.27f);                    /* cosmoparms.f90:30 */
static const double K_h = F32(0.7f);                          /* :28 */
static const double K_minlogtau = -20.0;                      /* radiation_tables.F90:45 */
static const double K_maxlogtau = 4.0;                        /* :46 */

static double K_ev2k(void) { return (double)(1.0f / 8.617e-05f); }            /* cgsconstants.f90:39 */
static double K_ev2fr(void) { return F32(0.241838e15f); }                      /* :53 */
static double K_temph0(void) { return K_eth0 * K_ev2k(); }                     /* :80 */
static double K_colh0(void) {                                                  /* :86 */
  /* 1.3e-8*fh0*xih0/(eth0*eth0): 1.3e-8 default real, fh0, xih0, eth0 real(dp) */
  const double fh0 = F32(0.83f), xih0 = F32(1.0f);
  return F32(1.3e-8f) * fh0 * xih0 / (K_eth0 * K_eth0);
}
static double K_sigma_HI(void) { return 1.0 * F32(6.30e-18f); }               /* cgsphotoconstants.f90:24 */
static double K_ion_freq_HI(void) { return K_ev2fr() * K_eth0; }               /* :31 */
static double K_ion_freq_HeII(void) { return K_ev2fr() * K_ethe1; }            /* :33 */
static double K_two_pi_over_c_square(void) { return F32(2.0f) * K_pi / (K_c * K_c); } /* cgsconstants.f90:61 */
static double K_dlogtau(void) { return (K_maxlogtau - K_minlogtau) / (double)(float)ORC_NUMTAU; } /* radiation_tables.F90:47 */
static double K_H0(void);
static double K_R_SOLAR(void) { return F32(6.9599e10f); }                      /* cgsastroconstants.f90:23 */
static double K_Mpc(void) { return F32(1e6f) * F32(3.086e18f); }               /* :29-31 */
static double K_H0(void) { return K_h * F32(100.0f) * F32(1e5f) / K_Mpc(); }   /* cosmoparms.f90:41 */

void orc_get_constants(orc_constants *c) {
  memset(c, 0, sizeof(*c));
  c->pi = K_pi;
  c->sigma_HI_at_ion_freq = K_sigma_HI();
  c->eth0 = K_eth0;
  c->ev2k = K_ev2k();
  c->temph0 = K_temph0();
  c->colh0 = K_colh0();
  c->ev2fr = K_ev2fr();
  c->ion_freq_HI = K_ion_freq_HI();
  c->ion_freq_HeII = K_ion_freq_HeII();
  c->bb_MaxFreq = K_ion_freq_HeII() * 10.00; /* sed_parameters.f90:36 */
  c->two_pi_over_c_square = K_two_pi_over_c_square();
  c->bh00 = K_bh00;
  c->albpow = K_albpow;
  c->hplanck = K_hplanck;
  c->k_B = K_k_B;
  c->c_light = K_c;
  c->m_p = K_m_p;
  c->sigma_SB = K_sigma_SB;
  c->abu_he = K_abu_he;
  c->abu_c = K_abu_c;
  c->mu = (F32(1.0f) - K_abu_he) + F32(4.0f) * K_abu_he; /* abundances.f90:32 */
  c->h = F32(0.7f);                                       /* cosmoparms.f90:28 */
  c->Omega0 = F32(0.27f);
  c->Omega_B = F32(0.044f);
  c->Mpc = K_Mpc();
  /* cosmoparms.f90:41-42: H0=h*100.0*1e5/Mpc ; rho_crit_0=3.0*H0*H0/(8.0*pi*G_grav) */
  c->H0 = c->h * F32(100.0f) * F32(1e5f) / c->Mpc;
  c->rho_crit_0 = F32(3.0f) * c->H0 * c->H0 / (F32(8.0f) * K_pi * 6.6732e-8);
  c->YEAR = F32(3.15576E+07f);
  c->R_SOLAR = K_R_SOLAR();
  c->epsilon = K_epsilon;
  c->convergence_fraction = K_convergence_fraction;
  c->minimum_fractional_change = K_minimum_fractional_change;
  c->minimum_fraction_of_atoms = K_minimum_fraction_of_atoms;
  c->loss_fraction = 1e-2; /* c2ray_parameters.f90:67 */
  c->max_coldensh = K_max_coldensh;
  c->tau_photo_limit = K_tau_photo_limit;
  c->sqrt3 = (double)sqrtf(3.0f); /* column_density.f90:52 */
  c->sqrt2 = (double)sqrtf(2.0f); /* :53 */
  c->minlogtau = K_minlogtau;
  c->dlogtau = K_dlogtau();
  c->xh_initial = F32(2e-4f); /* ionfractions_module.F90:49 */
  c->bb_Teff = F32(5.0e4f);   /* sed_parameters.f90:29 */
  c->bb_S_star = 1e48;        /* :31 */
}

/* ------------------------------------------------------------------------------------------ */
/* romberg.f90                                                                                 */
/* ------------------------------------------------------------------------------------------ */
#define MAXPOW 14
/* romw(0:2**maxpow, -1:maxpow); only the columns actually used are kept: [p+1][x] */
typedef struct {
  double *w[MAXPOW + 2];
} romberg_t;

/* romberg.f90:22-90 */
static void romberg_initialisation(romberg_t *R, int nmax) {
  double a[MAXPOW + 1], b[MAXPOW + 1];
  static double s[MAXPOW + 1][MAXPOW + 1];
  int pmax = (int)lround(log((double)nmax) / (double)logf(2.0f)); /* nint(log(real(nmax,dp))/log(2.0)) */
  for (int p = -1; p <= pmax; ++p) R->w[p + 1] = (double *)calloc((size_t)(1 << (p < 0 ? 0 : p)) + 1, sizeof(double));
  memset(s, 0, sizeof(s));
  for (int k = 1; k <= pmax; ++k) {
    /* b(k) = -1.0 / (4.0 ** k - 1.0)  -- evaluated entirely in default real */
    float four_k = 1.0f;
    for (int q = 0; q < k; ++q) four_k *= 4.0f;
    b[k] = (double)(-1.0f / (four_k - 1.0f));
    a[k] = -b[k] * (double)four_k; /* - b (k) * 4.0 ** k */
  }
  for (int i = 1; i <= pmax; ++i) s[i][0] = 0.0;
  for (int k = 0; k <= pmax; ++k) {
    s[k][0] = 1.0;
    for (int j = 1; j <= pmax; ++j)
      for (int i = pmax; i >= j; --i) s[i][j] = a[j] * s[i][j - 1] + b[j] * s[i - 1][j - 1];
    for (int i = k; i <= pmax; ++i)
      for (int j = 0; j <= (1 << k); ++j) {
        int x = (1 << (i - k)) * j;
        R->w[i + 1][x] = s[i][i] * (double)(1 << (i - k)) + R->w[i + 1][x];
      }
    s[k][0] = 0.0;
  }
  R->w[0][0] = 1.0; /* romw(0,-1) */
  for (int i = 0; i <= pmax; ++i) {
    R->w[i + 1][0] = 0.5 * R->w[i + 1][0];
    R->w[i + 1][1 << i] = 0.5 * R->w[i + 1][1 << i];
  }
}
static void romberg_free(romberg_t *R, int pmax) {
  for (int p = -1; p <= pmax; ++p) free(R->w[p + 1]);
}

/* ------------------------------------------------------------------------------------------ */
/* rad_ini: radiation_tables.F90:95-126 and callees                                             */
/* ------------------------------------------------------------------------------------------ */
static void rad_ini_impl(double *thick, double *thin, double *heat_thick, double *heat_thin, orc_rad_diag *diag);
void orc_rad_ini(double *thick, double *thin, orc_rad_diag *diag) { rad_ini_impl(thick, thin, NULL, NULL, diag); }
/* rad_ini with isothermal=.false.: also stellar_heat_thick_table / stellar_heat_thin_table
 * (fill_heating_integrands_HI radiation_tables.F90:471-509, make_heat_tables_HI :547-565) */
void orc_rad_ini_heat(double *thick, double *thin, double *heat_thick, double *heat_thin) {
  rad_ini_impl(thick, thin, heat_thick, heat_thin, NULL);
}
static void rad_ini_impl(double *thick, double *thin, double *heat_thick, double *heat_thin, orc_rad_diag *diag) {
  const int NF = ORC_NUMFREQ, NT = ORC_NUMTAU;
  const double pi = K_pi;
  const double ion_freq_HI = K_ion_freq_HI();
  const double tpocs = K_two_pi_over_c_square();

  /* spectrum_parms, case "B" (radiation_sed_parameters.F90:82-164) */
  double T_eff = fmax(fmin(F32(5.0e4f), F32(1e6f)), F32(2000.f));
  double S_star = 1e48;
  double MinFreq = ion_freq_HI;               /* bb_MinFreq, sed_parameters.f90:35 */
  double MaxFreq = K_ion_freq_HeII() * 10.00; /* bb_MaxFreq */
  double bb_luminosity_unscaled = K_sigma_SB * T_eff * T_eff * T_eff * T_eff;
  double R_star = K_R_SOLAR();
  double L_star = 4.0 * pi * R_star * R_star * bb_luminosity_unscaled;
  double h_over_kT = K_hplanck / (K_k_B * T_eff);
  /* freq_min_src is intent(out) but never assigned in spectrum_parms; setup_scalingfactors takes
   * max(ion_freq_HI, freq_min_src) (radiation_sizes.f90:61).  The restatement assumes the
   * undefined value does not exceed ion_freq_HI. */
  double freq_max_src = MaxFreq;

  /* setup_scalingfactors (radiation_sizes.f90:36-89) */
  double freq_max = freq_max_src;
  double freq_min = ion_freq_HI;
  double delta_freq = (freq_max - freq_min) / (double)(float)NF;
  const double pl_index_cross_section_HI = 2.8;

  /* romberg_initialisation (NumFreq) */
  romberg_t R;
  romberg_initialisation(&R, NF);
  const double *romw7 = R.w[7 + 1];

  /* spec_diag (radiation_sed_parameters.F90:172-224) with integrate_sed (:226-283), "B","S" */
  double S_star_unscaled;
  {
    double freq_step = (MaxFreq - MinFreq) / (double)(float)NF;
    double integral = 0.0;
    for (int i = 0; i <= NF; ++i) {
      double frequency = MinFreq + freq_step * (double)(float)i;
      double weight = freq_step;
      double integrand;
      if (frequency * h_over_kT <= 709.0)
        integrand = tpocs * frequency * frequency / (exp(frequency * h_over_kT) - 1.0);
      else
        integrand = tpocs * frequency * frequency / (exp((frequency * h_over_kT) / 2.0)) /
                    (exp((frequency * h_over_kT) / 2.0));
      /* scalar_romberg (romberg.f90:100-149), ny=0 => py=-1, romw(0,-1)=1 */
      integral = integral + integrand * weight * romw7[i] * R.w[0][0];
    }
    S_star_unscaled = F32(4.0f) * pi * R_star * R_star * integral;
  }
  double S_scaling = S_star / S_star_unscaled;
  R_star = sqrt(S_scaling) * R_star;
  L_star = S_scaling * L_star;
  (void)L_star;
  double R_star2 = R_star * R_star;

  /* spec_integration (radiation_tables.F90:130-236) */
  double *tau = (double *)malloc(sizeof(double) * (NT + 1));
  const double dlogtau = K_dlogtau();
  for (int i = 1; i <= NT; ++i) tau[i] = pow(F32(10.0f), K_minlogtau + dlogtau * (double)(float)(i - 1));
  tau[0] = 0.0;
  double frequency[ORC_NUMFREQ + 1], cs[ORC_NUMFREQ + 1], sed[ORC_NUMFREQ + 1];
  for (int i = 0; i <= NF; ++i) frequency[i] = freq_min + delta_freq * (double)(float)i; /* :264-272 */
  for (int i = 0; i <= NF; ++i) cs[i] = pow(frequency[i] / freq_min, -pl_index_cross_section_HI); /* :276-297 */
  for (int i = 0; i <= NF; ++i) { /* BB_SED :434-452 */
    if (frequency[i] * h_over_kT < F32(700.0f))
      sed[i] = 4.0 * pi * R_star2 * tpocs * frequency[i] * frequency[i] /
               (exp(frequency[i] * h_over_kT) - F32(1.0f));
    else
      sed[i] = 0.0;
  }
  for (int it = 0; it <= NT; ++it) { /* fill_photo_integrands :361-430 + make_photo_tables :524-543 */
    double a_thick = 0.0, a_thin = 0.0, h_thick = 0.0, h_thin = 0.0;
    for (int i = 0; i <= NF; ++i) {
      double f_thick, f_thin;
      if (tau[it] * cs[i] < F32(700.0f)) {
        f_thick = sed[i] * exp(-tau[it] * cs[i]);
        f_thin = sed[i] * cs[i] * exp(-tau[it] * cs[i]);
      } else {
        f_thick = 0.0;
        f_thin = 0.0;
      }
      /* vector_romberg romberg.f90:158-187 with vector_weight = delta_freq */
      a_thick = a_thick + f_thick * delta_freq * romw7[i];
      a_thin = a_thin + f_thin * delta_freq * romw7[i];
      /* heating integrands :482-487: hplanck*(frequency-ion_freq_HI)*photo integrand */
      h_thick = h_thick + (K_hplanck * (frequency[i] - ion_freq_HI) * f_thick) * delta_freq * romw7[i];
      h_thin = h_thin + (K_hplanck * (frequency[i] - ion_freq_HI) * f_thin) * delta_freq * romw7[i];
    }
    thick[it] = a_thick;
    thin[it] = a_thin;
    if (heat_thick) heat_thick[it] = h_thick;
    if (heat_thin) heat_thin[it] = h_thin;
  }
  if (diag) {
    diag->S_star_unscaled = S_star_unscaled;
    diag->S_scaling = S_scaling;
    diag->R_star = R_star;
    diag->h_over_kT = h_over_kT;
    diag->freq_min = freq_min;
    diag->freq_max = freq_max;
    diag->delta_freq = delta_freq;
    memcpy(diag->romw7, romw7, sizeof(double) * (NF + 1));
  }
  free(tau);
  romberg_free(&R, 7);
}

/* ------------------------------------------------------------------------------------------ */
/* state                                                                                        */
/* ------------------------------------------------------------------------------------------ */
struct orc_state {
  int mesh[3];
  size_t ncell;
  float *ndens;                  /* density_module.F90:22 */
  double *xh, *xh_av, *xh_intermed; /* ionfractions_module.F90:22, evolve_data.F90:50-58 */
  double *phih_grid, *coldensh_out; /* evolve_data.F90:40-48 */
  double dr[3], vol;             /* grid.F90:25,29 */
  int type_of_clumping;
  float clumping;                /* clumping_module.F90:17 (default real!) */
  float *clumping_grid;
  int use_LLS, type_of_LLS;
  double coldensh_LLS, R_max_LLS;
  float *LLS_grid;
  double temper_val;
  int NumSrc;
  int32_t *srcpos;
  double *NormFlux_stellar;
  double S_star;
  double thick[ORC_NUMTAU + 1], thin[ORC_NUMTAU + 1];
  double loss_fraction;
  int walk_order;
  int rank, npr, nthreads;
  int omp_in_source; /* 1: the threads share one source (evolve_source.F90:141-186), 0: one source per thread */
  /* photonstatistics module variables */
  double h0_before, h1_before, h0_after, h1_after, totrec, totcollisions, dh0, total_ion;
  double LLS_loss, photon_loss, grtotal_ion, grtotal_src;
  /* per-do_source working variables (evolve_data.F90:60-66, evolve_source.F90:44-48) */
  int last_l[3], last_r[3], lastpos_l[3], lastpos_r[3];
  double photon_loss_src_thread;
  int64_t updates;
  /* non-isothermal path (c2ray_parameters.f90:28 isothermal=.false.) */
  int isothermal;                     /* 1 by default, as shipped */
  int cosmological;                   /* c2ray_parameters.f90:105 */
  double zred;                        /* cosmology.F90:42, current redshift (cosmo_cool) */
  double heat_thick[ORC_NUMTAU + 1], heat_thin[ORC_NUMTAU + 1]; /* stellar_heat_*_table(:,1) */
  double *phiheat_grid;               /* evolve_data.F90:42 */
  float *temperature_grid;            /* temperature_module.F90:21-35: (current, average, intermed) per cell */
  double cie_cool[61], cool_mintemp, cool_dtemp; /* cooling.f90:26-29 */
  /* iteration dump (evolve.F90:285-324) kept in memory */
  int dump_at_iter, have_dump, dump_niter;
  double dump_photon_loss_all;
  double *dump_phih, *dump_xh_av, *dump_xh_intermed;
};

orc_state *orc_create(int m1, int m2, int m3) {
  orc_state *s = (orc_state *)calloc(1, sizeof(orc_state));
  s->mesh[0] = m1;
  s->mesh[1] = m2;
  s->mesh[2] = m3;
  s->ncell = (size_t)m1 * m2 * m3;
  s->ndens = (float *)calloc(s->ncell, sizeof(float));
  s->xh = (double *)calloc(s->ncell, sizeof(double));
  s->xh_av = (double *)calloc(s->ncell, sizeof(double));
  s->xh_intermed = (double *)calloc(s->ncell, sizeof(double));
  s->phih_grid = (double *)calloc(s->ncell, sizeof(double));
  s->coldensh_out = (double *)calloc(s->ncell, sizeof(double));
  s->type_of_clumping = 1;
  s->clumping = 1.0f;
  s->use_LLS = 0;
  s->type_of_LLS = 1;
  s->temper_val = F32(1e4f);
  s->loss_fraction = 1e-2;
  s->npr = 1;
  s->nthreads = 1;
  s->isothermal = 1;
  s->cosmological = 1;
  return s;
}
void orc_destroy(orc_state *s) {
  if (!s) return;
  free(s->ndens); free(s->xh); free(s->xh_av); free(s->xh_intermed); free(s->phih_grid);
  free(s->coldensh_out); free(s->clumping_grid); free(s->LLS_grid); free(s->srcpos);
  free(s->NormFlux_stellar);
  free(s->dump_phih); free(s->dump_xh_av); free(s->dump_xh_intermed);
  free(s->phiheat_grid); free(s->temperature_grid);
  free(s);
}
void orc_set_tables(orc_state *s, const double *thick, const double *thin) {
  memcpy(s->thick, thick, sizeof(s->thick));
  memcpy(s->thin, thin, sizeof(s->thin));
}
void orc_set_density(orc_state *s, const float *ndens) { memcpy(s->ndens, ndens, s->ncell * sizeof(float)); }
void orc_set_geometry(orc_state *s, const double dr[3], double vol) {
  s->dr[0] = dr[0]; s->dr[1] = dr[1]; s->dr[2] = dr[2];
  s->vol = vol;
}
void orc_set_clumping(orc_state *s, int type, float clumping, const float *grid) {
  s->type_of_clumping = type;
  s->clumping = clumping;
  free(s->clumping_grid);
  s->clumping_grid = NULL;
  if (grid) {
    s->clumping_grid = (float *)malloc(s->ncell * sizeof(float));
    memcpy(s->clumping_grid, grid, s->ncell * sizeof(float));
  }
}
/* deterministic_clumping, clumping_module.F90:327-363 (type_of_clumping 3): the quadratic fit in ndens/avg_dens,
 * evaluated left to right in real(dp) (:356-357) and stored as default real */
void orc_deterministic_clumping(orc_state *s, double p1, double p2, double p3, double avg_dens) {
  s->type_of_clumping = 3;
  if (!s->clumping_grid) s->clumping_grid = (float *)malloc(s->ncell * sizeof(float));
  for (size_t c = 0; c < s->ncell; ++c) {
    const double nd = (double)s->ndens[c];
    s->clumping_grid[c] = (float)(p1 * nd / avg_dens * nd / avg_dens + p2 * nd / avg_dens + p3);
  }
}
const float *orc_clumping_grid(const orc_state *s) { return s->clumping_grid; }
void orc_set_lls(orc_state *s, int use_LLS, int type_of_LLS, double coldensh_LLS, const float *grid,
                 double R_max_LLS) {
  s->use_LLS = use_LLS;
  s->type_of_LLS = type_of_LLS;
  s->coldensh_LLS = coldensh_LLS;
  s->R_max_LLS = R_max_LLS;
  free(s->LLS_grid);
  s->LLS_grid = NULL;
  if (grid) {
    s->LLS_grid = (float *)malloc(s->ncell * sizeof(float));
    memcpy(s->LLS_grid, grid, s->ncell * sizeof(float));
  }
}
void orc_set_temperature(orc_state *s, double t) { s->temper_val = t; }
void orc_set_sources(orc_state *s, int NumSrc, const int32_t *srcpos, const double *nf, double S_star) {
  free(s->srcpos);
  free(s->NormFlux_stellar);
  s->NumSrc = NumSrc;
  s->srcpos = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)(NumSrc > 0 ? NumSrc : 1));
  s->NormFlux_stellar = (double *)malloc(sizeof(double) * (size_t)(NumSrc > 0 ? NumSrc : 1));
  if (NumSrc > 0) {
    memcpy(s->srcpos, srcpos, sizeof(int32_t) * 3 * (size_t)NumSrc);
    memcpy(s->NormFlux_stellar, nf, sizeof(double) * (size_t)NumSrc);
  }
  s->S_star = S_star;
}
void orc_set_xh(orc_state *s, const double *xh) { memcpy(s->xh, xh, s->ncell * sizeof(double)); }
void orc_set_xh_av(orc_state *s, const double *x) { memcpy(s->xh_av, x, s->ncell * sizeof(double)); }
void orc_set_loss_fraction(orc_state *s, double lf) { s->loss_fraction = lf; }
void orc_set_walk_order(orc_state *s, int order) { s->walk_order = order; }
void orc_set_rank(orc_state *s, int rank, int npr) { s->rank = rank; s->npr = npr; }
void orc_set_threads(orc_state *s, int n) { s->nthreads = n < 1 ? 1 : n; }
/* how orc_set_threads(n > 1) is used by pass_all_sources: 0 = one source per thread with private rate grids (the MPI
 * picture, do_grid_static + MPI_ALLREDUCE), 1 = all threads inside one source (the OpenMP build,
 * evolve_source.F90:141-186: 6 axes, 12 planes, 8 octants) */
void orc_set_omp_in_source(orc_state *s, int on) { s->omp_in_source = on ? 1 : 0; }

/* ---- non-isothermal inputs ------------------------------------------------------------------ */
/* isothermal=.false. (c2ray_parameters.f90:28): allocates phiheat_grid (evolve_data.F90:77) and
 * temperature_grid, filled with temper_val (temperature_array_init temperature_module.F90:44-67) */
void orc_set_isothermal(orc_state *s, int isothermal) {
  s->isothermal = isothermal ? 1 : 0;
  if (!s->isothermal && !s->phiheat_grid) {
    s->phiheat_grid = (double *)calloc(s->ncell, sizeof(double));
    s->temperature_grid = (float *)malloc(3 * s->ncell * sizeof(float));
    for (size_t p = 0; p < 3 * s->ncell; ++p) s->temperature_grid[p] = (float)s->temper_val;
  }
}
void orc_set_heat_tables(orc_state *s, const double *heat_thick, const double *heat_thin) {
  memcpy(s->heat_thick, heat_thick, sizeof(s->heat_thick));
  memcpy(s->heat_thin, heat_thin, sizeof(s->heat_thin));
}
/* setup_cool cooling.f90:64-87: 61 rows (log10 T, log10 Lambda) of tables/corocool.tab */
void orc_set_cooling_table(orc_state *s, const double *log10_temp, const double *log10_cool) {
  s->cool_mintemp = log10_temp[0];
  s->cool_dtemp = log10_temp[1] - log10_temp[0];
  for (int i = 0; i < 61; ++i) s->cie_cool[i] = pow(10.0, log10_cool[i]);
}
void orc_set_redshift(orc_state *s, double zred, int cosmological) {
  s->zred = zred;
  s->cosmological = cosmological;
}
float *orc_temperature_grid(orc_state *s) { return s->temperature_grid; }
double *orc_phiheat(orc_state *s) { return s->phiheat_grid; }
double *orc_xh(orc_state *s) { return s->xh; }
double *orc_xh_av(orc_state *s) { return s->xh_av; }
double *orc_xh_intermed(orc_state *s) { return s->xh_intermed; }
double *orc_phih(orc_state *s) { return s->phih_grid; }
double *orc_coldensh_out(orc_state *s) { return s->coldensh_out; }

/* Fortran modulo(a,p) for p>0 */
static inline int f_modulo(int a, int p) {
  int r = a % p;
  return r < 0 ? r + p : r;
}
/* Fortran sign(1,x): +1 for x>=0 */
static inline int f_sign1(int x) { return x >= 0 ? 1 : -1; }
/* (i,j,k) 1-based -> linear, i fastest */
static inline size_t lin(const orc_state *s, int i, int j, int k) {
  return ((size_t)(k - 1) * s->mesh[1] + (size_t)(j - 1)) * s->mesh[0] + (size_t)(i - 1);
}

/* ------------------------------------------------------------------------------------------ */
/* column_density.f90                                                                           */
/* ------------------------------------------------------------------------------------------ */
/* column_density.f90:276-293 */
static inline double weightf(double cd) {
  double sig = K_sigma_HI();
  return F32(1.0f) / fmax(0.6, cd * sig);
}

/* column_density.f90:29-271 */
static void cinterp(const orc_state *s, const double *coldensh_out, const int pos[3],
                    const int srcpos[3], double *cdensi_out, double *path_out) {
  const double sqrt3 = (double)sqrtf(3.0f), sqrt2 = (double)sqrtf(2.0f);
  int i = pos[0], j = pos[1], k = pos[2];
  int i0 = srcpos[0], j0 = srcpos[1], k0 = srcpos[2];
  int idel = i - i0, jdel = j - j0, kdel = k - k0;
  int idela = abs(idel), jdela = abs(jdel), kdela = abs(kdel);
  int sgni = f_sign1(idel), sgnj = f_sign1(jdel), sgnk = f_sign1(kdel);
  int im = i - sgni, jm = j - sgnj, km = k - sgnk;
  double di = (double)(float)idel, dj = (double)(float)jdel, dk = (double)(float)kdel;
  double alam, xc, yc, zc, dx, dy, dz, s1, s2, s3, s4, c1, c2, c3, c4, w1, w2, w3, w4;
  double cdensi = 0.0, path = 0.0;
  const int *mesh = s->mesh;
  if (kdela >= jdela && kdela >= idela) { /* :108 */
    alam = (double)((float)(km - k0) + (float)sgnk * 0.5f) / dk;
    xc = alam * di + (double)(float)i0;
    yc = alam * dj + (double)(float)j0;
    dx = F32(2.0f) * fabs(xc - (double)((float)im + 0.5f * (float)sgni));
    dy = F32(2.0f) * fabs(yc - (double)((float)jm + 0.5f * (float)sgnj));
    s1 = (1.0 - dx) * (1.0 - dy);
    s2 = (1.0 - dy) * dx;
    s3 = (1.0 - dx) * dy;
    s4 = dx * dy;
    int ip = f_modulo(i - 1, mesh[0]) + 1, imp = f_modulo(im - 1, mesh[0]) + 1;
    int jp = f_modulo(j - 1, mesh[1]) + 1, jmp = f_modulo(jm - 1, mesh[1]) + 1;
    int kmp = f_modulo(km - 1, mesh[2]) + 1;
    c1 = coldensh_out[lin(s, imp, jmp, kmp)];
    c2 = coldensh_out[lin(s, ip, jmp, kmp)];
    c3 = coldensh_out[lin(s, imp, jp, kmp)];
    c4 = coldensh_out[lin(s, ip, jp, kmp)];
    w1 = s1 * weightf(c1);
    w2 = s2 * weightf(c2);
    w3 = s3 * weightf(c3);
    w4 = s4 * weightf(c4);
    cdensi = (c1 * w1 + c2 * w2 + c3 * w3 + c4 * w4) / (w1 + w2 + w3 + w4);
    if (kdela == 1 && (idela == 1 || jdela == 1)) { /* :152-158 */
      if (idela == 1 && jdela == 1) cdensi = sqrt3 * cdensi;
      else cdensi = sqrt2 * cdensi;
    }
    path = sqrt((di * di + dj * dj) / (dk * dk) + F32(1.0f));
  } else if (jdela >= idela && jdela >= kdela) { /* :173 */
    alam = (double)((float)(jm - j0) + (float)sgnj * 0.5f) / dj;
    zc = alam * dk + (double)(float)k0;
    xc = alam * di + (double)(float)i0;
    dz = F32(2.0f) * fabs(zc - (double)((float)km + 0.5f * (float)sgnk));
    dx = F32(2.0f) * fabs(xc - (double)((float)im + 0.5f * (float)sgni));
    s1 = (1.0 - dx) * (1.0 - dz);
    s2 = (1.0 - dz) * dx;
    s3 = (1.0 - dx) * dz;
    s4 = dx * dz;
    int ip = f_modulo(i - 1, mesh[0]) + 1, imp = f_modulo(im - 1, mesh[0]) + 1;
    int jmp = f_modulo(jm - 1, mesh[1]) + 1;
    int kp = f_modulo(k - 1, mesh[2]) + 1, kmp = f_modulo(km - 1, mesh[2]) + 1;
    c1 = coldensh_out[lin(s, imp, jmp, kmp)];
    c2 = coldensh_out[lin(s, ip, jmp, kmp)];
    c3 = coldensh_out[lin(s, imp, jmp, kp)];
    c4 = coldensh_out[lin(s, ip, jmp, kp)];
    w1 = s1 * weightf(c1);
    w2 = s2 * weightf(c2);
    w3 = s3 * weightf(c3);
    w4 = s4 * weightf(c4);
    cdensi = (c1 * w1 + c2 * w2 + c3 * w3 + c4 * w4) / (w1 + w2 + w3 + w4);
    if (jdela == 1 && (idela == 1 || kdela == 1)) {
      if (idela == 1 && kdela == 1) cdensi = sqrt3 * cdensi;
      else cdensi = sqrt2 * cdensi;
    }
    path = sqrt((di * di + dk * dk) / (dj * dj) + F32(1.0f));
  } else if (idela >= jdela && idela >= kdela) { /* :226 */
    alam = (double)((float)(im - i0) + (float)sgni * 0.5f) / di;
    zc = alam * dk + (double)(float)k0;
    yc = alam * dj + (double)(float)j0;
    dz = F32(2.0f) * fabs(zc - (double)((float)km + 0.5f * (float)sgnk));
    dy = F32(2.0f) * fabs(yc - (double)((float)jm + 0.5f * (float)sgnj));
    s1 = (1.0 - dz) * (1.0 - dy);
    s2 = (1.0 - dz) * dy;
    s3 = (1.0 - dy) * dz;
    s4 = dy * dz;
    int imp = f_modulo(im - 1, mesh[0]) + 1;
    int jp = f_modulo(j - 1, mesh[1]) + 1, jmp = f_modulo(jm - 1, mesh[1]) + 1;
    int kp = f_modulo(k - 1, mesh[2]) + 1, kmp = f_modulo(km - 1, mesh[2]) + 1;
    c1 = coldensh_out[lin(s, imp, jmp, kmp)];
    c2 = coldensh_out[lin(s, imp, jp, kmp)];
    c3 = coldensh_out[lin(s, imp, jmp, kp)];
    c4 = coldensh_out[lin(s, imp, jp, kp)];
    w1 = s1 * weightf(c1);
    w2 = s2 * weightf(c2);
    w3 = s3 * weightf(c3);
    w4 = s4 * weightf(c4);
    cdensi = (c1 * w1 + c2 * w2 + c3 * w3 + c4 * w4) / (w1 + w2 + w3 + w4);
    if (idela == 1 && (jdela == 1 || kdela == 1)) {
      if (jdela == 1 && kdela == 1) cdensi = sqrt3 * cdensi;
      else cdensi = sqrt2 * cdensi;
    }
    path = sqrt(F32(1.0f) + (dj * dj + dk * dk) / (di * di));
  }
  *cdensi_out = cdensi;
  *path_out = path;
}

void orc_cinterp(const orc_state *s, const int pos[3], const int srcpos[3], double *cdensi, double *path) {
  cinterp(s, s->coldensh_out, pos, srcpos, cdensi, path);
}

/* ------------------------------------------------------------------------------------------ */
/* radiation_photoionrates.F90                                                                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double tau, odpos, residual; int ipos, ipos_p1; } tablepos;

/* radiation_photoionrates.F90:184-208 (NumFreqBnd = 1) */
static inline tablepos set_tau_table_positions(double tau) {
  tablepos p;
  p.tau = log10(fmax(1.0e-20, tau));
  p.odpos = fmin((double)ORC_NUMTAU, fmax(0.0, F32(1.0f) + (p.tau - K_minlogtau) / K_dlogtau()));
  p.ipos = (int)p.odpos;
  p.residual = p.odpos - (double)p.ipos;
  p.ipos_p1 = p.ipos + 1 < ORC_NUMTAU ? p.ipos + 1 : ORC_NUMTAU;
  return p;
}
/* :212-228 */
static inline double read_table(const double *table, const tablepos *p) {
  return table[p->ipos] + (table[p->ipos_p1] - table[p->ipos]) * p->residual;
}

typedef struct { double photo_cell_HI, photo_in, photo_out, heat; } photrates;

/* photoion_rates :71-179 with photo_lookuptable :233-317, table_type "B", isothermal */
static inline photrates photoion_rates(const orc_state *s, double colum_in_HI, double colum_out_HI,
                                       double vol, double NFlux) {
  photrates phi = {0.0, 0.0, 0.0, 0.0}; /* set_photrates_to_zero :442-450 */
  const double sigma_HI = K_sigma_HI(); /* radiation_sizes.f90:77 */
  double tau_in = colum_in_HI * sigma_HI;
  double tau_out = colum_out_HI * sigma_HI;
  tablepos pin = set_tau_table_positions(tau_in);
  tablepos pout = set_tau_table_positions(tau_out);
  if (NFlux > 0.0) { /* :118 */
    double phi_photo_in_all = NFlux * read_table(s->thick, &pin);
    double phi_photo_out_all, phi_photo_all;
    if (fabs(tau_out - tau_in) > K_tau_photo_limit) {
      phi_photo_out_all = NFlux * read_table(s->thick, &pout);
      phi_photo_all = phi_photo_in_all - phi_photo_out_all;
    } else {
      phi_photo_all = NFlux * (tau_out - tau_in) * read_table(s->thin, &pin);
      phi_photo_out_all = phi_photo_in_all - phi_photo_all;
    }
    /* phi = phi + lookup  (photrates_add :421-436); 0.0 + x */
    phi.photo_in = 0.0 + (0.0 + phi_photo_in_all);
    phi.photo_out = 0.0 + (0.0 + phi_photo_out_all);
    phi.photo_cell_HI = 0.0 + (0.0 + phi_photo_all / vol);
  }
  if (!s->isothermal && NFlux > 0.0) { /* :140-165 -> heat_lookuptable :323-417, table_type "B", one sub-band */
    double tau_cell_HI = (colum_out_HI - colum_in_HI) * sigma_HI; /* :144-146 */
    double phi_heat_in_HI = NFlux * read_table(s->heat_thick, &pin);
    double phi_heat_HI;
    if (fabs(tau_out - tau_in) > K_tau_heat_limit) {
      double phi_heat_out_HI = NFlux * read_table(s->heat_thick, &pout);
      phi_heat_HI = (phi_heat_in_HI - phi_heat_out_HI) / vol;
    } else {
      phi_heat_HI = NFlux * tau_cell_HI * read_table(s->heat_thin, &pin);
      phi_heat_HI = phi_heat_HI / vol;
    }
    phi.heat = 0.0 + (0.0 + phi_heat_HI); /* f_heat = f_heat + df_heat ; phi = phi + ... */
  }
  return phi;
}

void orc_photoion_rates(const orc_state *s, double colum_in, double colum_out, double vol,
                        double normflux, double out[3]) {
  photrates p = photoion_rates(s, colum_in, colum_out, vol, normflux);
  out[0] = p.photo_cell_HI;
  out[1] = p.photo_in;
  out[2] = p.photo_out;
}

/* ------------------------------------------------------------------------------------------ */
/* evolve_point.F90:83-299  evolve0D                                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const orc_state *s;
  double *coldensh_out; /* may be thread-private */
  double *phih_grid;    /* may be thread-private */
  double *phiheat_grid; /* may be thread-private; NULL when isothermal */
  int last_l[3], last_r[3];
  double photon_loss_src_thread;
  int64_t updates;
} walk_ctx;

static void evolve0D(walk_ctx *w, const int rtpos[3], int ns) {
  const orc_state *s = w->s;
  const int *src = &s->srcpos[3 * (ns - 1)];
  int pos[3];
  for (int d = 0; d < 3; ++d) pos[d] = f_modulo(rtpos[d] - 1, s->mesh[d]) + 1; /* :122-124 */
  size_t p = lin(s, pos[0], pos[1], pos[2]);
  if (w->coldensh_out[p] == 0.0) { /* :128 */
    int stop_rad_transfer = 0;
    double h_av1 = fmax(s->xh_av[p], K_epsilon);                /* :137 */
    double h_av0 = fmax(F32(1.0f) - h_av1, K_epsilon);          /* :140 */
    double ndens_p = (double)s->ndens[p];                       /* :145 */
    double coldensh_in, path, vol_ph;
    w->updates++;
    if (rtpos[0] == src[0] && rtpos[1] == src[1] && rtpos[2] == src[2]) { /* :151-160 */
      coldensh_in = 0.0;
      path = F32(0.5f) * s->dr[0];
      vol_ph = s->dr[0] * s->dr[1] * s->dr[2];
    } else {
      int srcp[3] = {src[0], src[1], src[2]};
      cinterp(s, w->coldensh_out, rtpos, srcp, &coldensh_in, &path); /* :165 */
      path = path * s->dr[0];
      double xs = s->dr[0] * (double)(float)(rtpos[0] - src[0]);
      double ys = s->dr[1] * (double)(float)(rtpos[1] - src[1]);
      double zs = s->dr[2] * (double)(float)(rtpos[2] - src[2]);
      double dist2 = xs * xs + ys * ys + zs * zs;
      vol_ph = F32(4.0f) * K_pi * dist2 * path; /* :177 */
      if (s->use_LLS) { /* :186-196 */
        if (s->type_of_LLS == 3) {
          if (dist2 > s->R_max_LLS * s->R_max_LLS) stop_rad_transfer = 1;
        } else {
          double coldensh_LLS = s->coldensh_LLS;
          if (s->type_of_LLS == 2) coldensh_LLS = (double)s->LLS_grid[p]; /* LLS_point LLS.F90:199-210 */
          coldensh_in = coldensh_in + coldensh_LLS * path / s->dr[0];
        }
      }
    }
    if (coldensh_in > K_max_coldensh) stop_rad_transfer = 1; /* :201 */
    /* niter == -1 branch (:213-245) is dead: niter >= 1 always */
    double cd_out = coldensh_in + h_av0 * ndens_p * path; /* :247-248, coldens doric.f90:141-155 */
    w->coldensh_out[p] = cd_out;
    photrates phi;
    if (!stop_rad_transfer) { /* :254-270 */
      phi = photoion_rates(s, coldensh_in, cd_out, vol_ph, s->NormFlux_stellar[ns - 1]);
      phi.photo_cell_HI = phi.photo_cell_HI / (h_av0 * ndens_p);
      /* total_LLS_loss(phi%photo_in_HI*vol/vol_ph, ...): photo_in_HI is never assigned (stays 0)
       * so LLS_loss += 0*(1-exp(-tau_LLS)) == 0. */
    } else {
      phi.photo_cell_HI = 0.0;
      phi.photo_in = 0.0;
      phi.photo_out = 0.0;
      phi.heat = 0.0;
    }
    w->phih_grid[p] = w->phih_grid[p] + phi.photo_cell_HI; /* :283-284 */
    if (!s->isothermal) w->phiheat_grid[p] = w->phiheat_grid[p] + phi.heat; /* :285-286 */
    if (rtpos[0] == w->last_l[0] || rtpos[1] == w->last_l[1] || rtpos[2] == w->last_l[2] ||
        rtpos[0] == w->last_r[0] || rtpos[1] == w->last_r[1] || rtpos[2] == w->last_r[2]) { /* :290-295 */
      w->photon_loss_src_thread = w->photon_loss_src_thread + phi.photo_out * s->vol / vol_ph;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* evolve_source.F90                                                                            */
/* ------------------------------------------------------------------------------------------ */
/* evolve2D :227-267 */
static void evolve2D(walk_ctx *w, int rtpos[3], int ns) {
  const int *src = &w->s->srcpos[3 * (ns - 1)];
  for (int j = src[1]; j <= w->last_r[1]; ++j) {
    rtpos[1] = j;
    for (int i = src[0]; i <= w->last_r[0]; ++i) { rtpos[0] = i; evolve0D(w, rtpos, ns); }
    for (int i = src[0] - 1; i >= w->last_l[0]; --i) { rtpos[0] = i; evolve0D(w, rtpos, ns); }
  }
  for (int j = src[1] - 1; j >= w->last_l[1]; --j) {
    rtpos[1] = j;
    for (int i = src[0]; i <= w->last_r[0]; ++i) { rtpos[0] = i; evolve0D(w, rtpos, ns); }
    for (int i = src[0] - 1; i >= w->last_l[0]; --i) { rtpos[0] = i; evolve0D(w, rtpos, ns); }
  }
}

/* One directed half-line / quarter-plane / octant sweep.  sg[d] = +1: src+1..last_r, -1:
 * src-1..last_l(-1), 0: fixed at src.  Loop nest k outer, j, i inner as in evolve1D_axis
 * :273-330, evolve2D_plane :336-473, evolve3D_quadrant :479-591. */
static void sweep_signed(walk_ctx *w, int ns, const int sg[3]) {
  const int *src = &w->s->srcpos[3 * (ns - 1)];
  int lo[3], hi[3], st[3];
  for (int d = 0; d < 3; ++d) {
    if (sg[d] > 0) { lo[d] = src[d] + 1; hi[d] = w->last_r[d]; st[d] = 1; }
    else if (sg[d] < 0) { lo[d] = src[d] - 1; hi[d] = w->last_l[d]; st[d] = -1; }
    else { lo[d] = src[d]; hi[d] = src[d]; st[d] = 1; }
  }
  int rtpos[3];
  for (int k = lo[2]; st[2] > 0 ? k <= hi[2] : k >= hi[2]; k += st[2]) {
    rtpos[2] = k;
    for (int j = lo[1]; st[1] > 0 ? j <= hi[1] : j >= hi[1]; j += st[1]) {
      rtpos[1] = j;
      for (int i = lo[0]; st[0] > 0 ? i <= hi[0] : i >= hi[0]; i += st[0]) {
        rtpos[0] = i;
        evolve0D(w, rtpos, ns);
      }
    }
  }
}

/* Chebyshev-shell walk with a scrambled order inside every shell (test pin for SURVEY A3) */
static void sweep_shells_scrambled(walk_ctx *w, int ns) {
  const int *src = &w->s->srcpos[3 * (ns - 1)];
  int rmax = 0;
  for (int d = 0; d < 3; ++d) {
    if (w->last_r[d] - src[d] > rmax) rmax = w->last_r[d] - src[d];
    if (src[d] - w->last_l[d] > rmax) rmax = src[d] - w->last_l[d];
  }
  for (int r = 0; r <= rmax; ++r) {
    size_t cap = (size_t)(2 * r + 1) * (2 * r + 1) * 6 + 8, n = 0;
    int *cells = (int *)malloc(cap * 3 * sizeof(int));
    for (int dk = -r; dk <= r; ++dk)
      for (int dj = -r; dj <= r; ++dj)
        for (int di = -r; di <= r; ++di) {
          int m = abs(di) > abs(dj) ? abs(di) : abs(dj);
          if (abs(dk) > m) m = abs(dk);
          if (m != r) continue;
          int p[3] = {src[0] + di, src[1] + dj, src[2] + dk};
          int ok = 1;
          for (int d = 0; d < 3; ++d) if (p[d] < w->last_l[d] || p[d] > w->last_r[d]) ok = 0;
          if (!ok) continue;
          cells[3 * n] = p[0]; cells[3 * n + 1] = p[1]; cells[3 * n + 2] = p[2];
          ++n;
        }
    /* deterministic scramble: visit in stride order with a stride coprime to n */
    size_t stride = 1;
    if (n > 2) { stride = n / 2 + 1; while (1) { size_t a = stride, b = n; while (b) { size_t t = a % b; a = b; b = t; } if (a == 1) break; ++stride; } }
    for (size_t q = 0, idx = (n > 0 ? (7 * (size_t)r) % n : 0); q < n; ++q, idx = (idx + stride) % n) {
      int rtpos[3] = {cells[3 * idx], cells[3 * idx + 1], cells[3 * idx + 2]};
      evolve0D(w, rtpos, ns);
    }
    free(cells);
  }
}

/* do_source evolve_source.F90:58-221 */
static void do_source(const orc_state *s, double *coldensh_out, double *phih_grid, double *phiheat_grid, int ns,
                      orc_source_report *rep) {
  walk_ctx w;
  w.s = s;
  w.coldensh_out = coldensh_out;
  w.phih_grid = phih_grid;
  w.phiheat_grid = phiheat_grid;
  w.updates = 0;
  const int *src = &s->srcpos[3 * (ns - 1)];
  int lastpos_l[3], lastpos_r[3];
  memset(coldensh_out, 0, s->ncell * sizeof(double)); /* :91  coldensh_out(:,:,:)=0.0 */
  for (int d = 0; d < 3; ++d) { /* :100-102, periodic_bc = .true. */
    int a = s->mesh[d] / 2 - 1 + s->mesh[d] % 2;
    int b = s->mesh[d] / 2;
    lastpos_r[d] = src[d] + (K_max_subbox < a ? K_max_subbox : a);
    lastpos_l[d] = src[d] - (K_max_subbox < b ? K_max_subbox : b);
  }
  int nbox = 0;
  double total_source_flux = s->NormFlux_stellar[ns - 1] * s->S_star; /* :119 */
  double photon_loss_src = total_source_flux;                         /* :121 */
  for (int d = 0; d < 3; ++d) { w.last_r[d] = src[d]; w.last_l[d] = src[d]; }
  while (photon_loss_src > s->loss_fraction * total_source_flux && w.last_r[2] < lastpos_r[2] &&
         w.last_l[2] > lastpos_l[2]) { /* :128-131 */
    nbox = nbox + 1;
    photon_loss_src = 0.0;
    w.photon_loss_src_thread = 0.0;
    for (int d = 0; d < 3; ++d) { /* :135-136 */
      int r = src[d] + K_subboxsize * nbox, l = src[d] - K_subboxsize * nbox;
      w.last_r[d] = r < lastpos_r[d] ? r : lastpos_r[d];
      w.last_l[d] = l > lastpos_l[d] ? l : lastpos_l[d];
    }
    if (s->omp_in_source && s->nthreads > 1) {
      /* OpenMP branch :141-186 executed by nthreads threads: the source cell, then the 6 axes, the 12 planes and
       * the 8 octants, each group shared between the threads with a barrier after it.  The sweeps of a group
       * touch disjoint cells; every sweep adds its boundary loss to a private counter
       * (photon_loss_src_thread(tn)), summed afterwards in sweep order. */
      if (nbox == 1) { int rtpos[3] = {src[0], src[1], src[2]}; evolve0D(&w, rtpos, ns); }
      static const int A[6][3] = {{1,0,0},{-1,0,0},{0,1,0},{0,-1,0},{0,0,1},{0,0,-1}};
      static const int P[12][3] = {{1,1,0},{1,-1,0},{-1,1,0},{-1,-1,0},{1,0,1},{-1,0,1},{-1,0,-1},{1,0,-1},
                                   {0,1,1},{0,-1,1},{0,1,-1},{0,-1,-1}};
      static const int Q[8][3] = {{1,1,1},{-1,1,1},{1,-1,1},{-1,-1,1},{1,1,-1},{-1,1,-1},{1,-1,-1},{-1,-1,-1}};
      const int (*groups[3])[3] = {A, P, Q};
      const int count[3] = {6, 12, 8};
      double part_loss[12];
      int64_t part_upd[12];
      for (int g = 0; g < 3; ++g) {
#pragma omp parallel for num_threads(s->nthreads) schedule(dynamic, 1)
        for (int n = 0; n < count[g]; ++n) {
          walk_ctx wl = w;
          wl.photon_loss_src_thread = 0.0;
          wl.updates = 0;
          sweep_signed(&wl, ns, groups[g][n]);
          part_loss[n] = wl.photon_loss_src_thread;
          part_upd[n] = wl.updates;
        }
        for (int n = 0; n < count[g]; ++n) {
          w.photon_loss_src_thread = w.photon_loss_src_thread + part_loss[n];
          w.updates += part_upd[n];
        }
      }
      photon_loss_src = photon_loss_src + w.photon_loss_src_thread;
    } else if (s->walk_order == 1) { /* OpenMP branch :141-186 executed by one thread */
      if (nbox == 1) { int rtpos[3] = {src[0], src[1], src[2]}; evolve0D(&w, rtpos, ns); }
      /* axes 1..6: +i,-i,+j,-j,+k,-k */
      for (int d = 0; d < 3; ++d) for (int sg = 1; sg >= -1; sg -= 2) {
        int v[3] = {0, 0, 0}; v[d] = sg; sweep_signed(&w, ns, v);
      }
      /* planes 1..12 (order of the reference's case list: z-fixed 4, y-fixed 4, x-fixed 4) */
      { const int P[12][3] = {{1,1,0},{1,-1,0},{-1,1,0},{-1,-1,0},{1,0,1},{-1,0,1},{-1,0,-1},{1,0,-1},
                              {0,1,1},{0,-1,1},{0,1,-1},{0,-1,-1}};
        for (int n = 0; n < 12; ++n) sweep_signed(&w, ns, P[n]); }
      { const int Q[8][3] = {{1,1,1},{-1,1,1},{1,-1,1},{-1,-1,1},{1,1,-1},{-1,1,-1},{1,-1,-1},{-1,-1,-1}};
        for (int n = 0; n < 8; ++n) sweep_signed(&w, ns, Q[n]); }
      photon_loss_src = photon_loss_src + w.photon_loss_src_thread;
    } else if (s->walk_order == 2) {
      sweep_shells_scrambled(&w, ns);
      photon_loss_src = w.photon_loss_src_thread;
    } else { /* serial branch :189-208 */
      int rtpos[3];
      for (int k = src[2]; k <= w.last_r[2]; ++k) { rtpos[2] = k; evolve2D(&w, rtpos, ns); }
      for (int k = src[2] - 1; k >= w.last_l[2]; --k) { rtpos[2] = k; evolve2D(&w, rtpos, ns); }
      photon_loss_src = w.photon_loss_src_thread;
    }
  }
  rep->nbox = nbox;
  rep->photon_loss_src = photon_loss_src;
  rep->updates = w.updates;
}

void orc_do_source(orc_state *s, int ns, orc_source_report *rep) {
  do_source(s, s->coldensh_out, s->phih_grid, s->phiheat_grid, ns, rep);
  s->photon_loss = s->photon_loss + rep->photon_loss_src; /* :216 */
}

/* evolve.F90:430-440 */
void orc_set_rates_to_zero(orc_state *s) {
  memset(s->phih_grid, 0, s->ncell * sizeof(double));
  if (!s->isothermal) memset(s->phiheat_grid, 0, s->ncell * sizeof(double)); /* :435 */
  s->photon_loss = 0.0;
  s->LLS_loss = 0.0;
}

/* pass_all_sources evolve.F90:444-495 -> do_grid_static master_slave.F90:74-96.
 * nthreads>1 emulates npr MPI ranks on host threads: private phih/coldensh_out per thread,
 * then the MPI_ALLREDUCE(SUM) of evolve.F90:599-602 in rank order. */
void orc_pass_all_sources(orc_state *s, orc_pass_report *rep) {
  int64_t sum_nbox = 0, updates = 0;
  double photon_loss = 0.0;
  const int mine = (s->NumSrc - s->rank + s->npr - 1) / s->npr; /* sources of this rank */
  if (s->nthreads <= 1 || mine <= 1 || s->omp_in_source) {
    for (int ns1 = 1 + s->rank; ns1 <= s->NumSrc; ns1 += s->npr) {
      orc_source_report r;
      do_source(s, s->coldensh_out, s->phih_grid, s->phiheat_grid, ns1, &r);
      photon_loss = photon_loss + r.photon_loss_src; /* evolve_source.F90:216 */
      sum_nbox += r.nbox;                            /* :219 */
      updates += r.updates;
    }
  } else {
    int T = s->nthreads < mine ? s->nthreads : mine;
    double **priv = (double **)calloc((size_t)T, sizeof(double *));
    double **privh = (double **)calloc((size_t)T, sizeof(double *));   /* phiheat_grid of the emulated rank */
    double *ploss = (double *)calloc((size_t)T, sizeof(double));
    int64_t *pnbox = (int64_t *)calloc((size_t)T, sizeof(int64_t));
    int64_t *pupd = (int64_t *)calloc((size_t)T, sizeof(int64_t));
#ifdef _OPENMP
#pragma omp parallel num_threads(T)
#endif
    {
#ifdef _OPENMP
      int t = omp_get_thread_num();
#else
      int t = 0;
#endif
      double *cd = (double *)malloc(s->ncell * sizeof(double));
      double *ph = (double *)calloc(s->ncell, sizeof(double));
      double *phh = s->isothermal ? NULL : (double *)calloc(s->ncell, sizeof(double));
      priv[t] = ph;
      privh[t] = phh;
      /* sources of this rank, dealt round-robin to the emulated ranks */
      int cnt = 0;
      for (int ns1 = 1 + s->rank; ns1 <= s->NumSrc; ns1 += s->npr, ++cnt) {
        if (cnt % T != t) continue;
        orc_source_report r;
        do_source(s, cd, ph, phh, ns1, &r);
        ploss[t] += r.photon_loss_src;
        pnbox[t] += r.nbox;
        pupd[t] += r.updates;
      }
      free(cd);
    }
#ifndef _OPENMP
    T = 1;
#endif
    for (int t = 0; t < T; ++t) {
      if (!priv[t]) continue;
      for (size_t p = 0; p < s->ncell; ++p) s->phih_grid[p] += priv[t][p];
      if (privh[t]) { /* evolve.F90:604-609 */
        for (size_t p = 0; p < s->ncell; ++p) s->phiheat_grid[p] += privh[t][p];
        free(privh[t]);
      }
      photon_loss += ploss[t];
      sum_nbox += pnbox[t];
      updates += pupd[t];
      free(priv[t]);
    }
    free(priv); free(privh); free(ploss); free(pnbox); free(pupd);
  }
  s->photon_loss = s->photon_loss + photon_loss;
  rep->photon_loss_all = s->photon_loss; /* photon_loss_all(:)=photon_loss(:) :485 */
  rep->sum_nbox_all = sum_nbox;
  rep->updates = updates;
}

/* ------------------------------------------------------------------------------------------ */
/* doric.f90:33-134, tped.f90:75-83                                                             */
/* ------------------------------------------------------------------------------------------ */
static inline double electrondens(double ndens, const double xh[2]) { return ndens * (xh[1] + K_abu_c); }

static inline void doric(double dt, double temp0, double rhe, double rhh, double xfh[2],
                         double xfh_av[2], double phih, float clumping) {
  (void)rhh;
  double brech0 = (double)clumping * K_bh00 * pow(temp0 / F32(1e4f), K_albpow);
  double sqrtt0 = sqrt(temp0);
  double acolh0 = K_colh0() * sqrtt0 * exp(-K_temph0() / temp0);
  double aphoth0 = phih;
  double xfh1old = xfh[1], xfh0old = xfh[0];
  double aih0 = aphoth0 + rhe * acolh0;
  double delth = aih0 + rhe * brech0;
  double eqxfh1 = aih0 / delth;
  double eqxfh0 = rhe * brech0 / delth;
  double deltht = delth * dt;
  double ee = exp(-deltht);
  xfh[1] = (xfh1old - eqxfh1) * ee + eqxfh1;
  xfh[0] = (xfh0old - eqxfh0) * ee + eqxfh0;
  if (xfh[0] < K_epsilon) { xfh[0] = K_epsilon; xfh[1] = 1.0 - K_epsilon; }
  double avg_factor;
  if (deltht < F32(1.0e-8f)) avg_factor = 1.0;
  else avg_factor = (F32(1.0f) - ee) / deltht;
  xfh_av[1] = eqxfh1 + (xfh1old - eqxfh1) * avg_factor;
  xfh_av[0] = 1.0 - xfh_av[1];
  if (xfh_av[0] < K_epsilon) xfh_av[0] = K_epsilon;
}

void orc_doric(const orc_state *s, double dt, double temp0, double rhe, double rhh, double xfh[2],
               double xfh_av[2], double phih, float clumping) {
  (void)s;
  doric(dt, temp0, rhe, rhh, xfh, xfh_av, phih, clumping);
}

/* coolin cooling.f90:38-59 */
static inline double coolin(const orc_state *s, double nucldens, double eldens, double temp0) {
  double tpos = (log10(temp0) - s->cool_mintemp) / s->cool_dtemp + 1.0;
  int itpos = (int)tpos;
  if (itpos < 1) itpos = 1;
  if (itpos > 61 - 1) itpos = 61 - 1;
  double dtpos = tpos - (double)(float)itpos;
  int itpos1 = itpos + 1 < 61 ? itpos + 1 : 61;
  return nucldens * eldens * (s->cie_cool[itpos - 1] + (s->cie_cool[itpos1 - 1] - s->cie_cool[itpos - 1]) * dtpos);
}

/* cosmo_cool cosmology.F90:198-225 */
static inline double cosmo_cool(const orc_state *s, double e_int) {
  double zred = s->zred;
  double dzdt = K_H0() * (F32(1.f) + zred) * sqrt(K_Omega0 * pow(F32(1.f) + zred, 3) + F32(1.f) - K_Omega0);
  return e_int * F32(2.0f) / (F32(1.0f) + zred) * dzdt;
}

/* thermal thermal.f90:22-176; final/average are left untouched when initial_temperature <= minitemp */
static void thermal(const orc_state *s, double dt, double initial_temperature, double *final_temperature,
                    double *average_temperature, double ndens_electron, double ndens_atom, const double h[2],
                    const double h_old[2], const double h_av[2], double heat) {
  /* temper2pressr tped.f90:41-53: (ndens+eldens)*k_B*temper */
  double internal_energy =
      (ndens_atom + electrondens(ndens_atom, h_old)) * K_k_B * initial_temperature / K_gamma1;
  double heating = heat;
  double cosmo_cool_rate = s->cosmological ? cosmo_cool(s, internal_energy) : 0.0;
  if (initial_temperature > K_minitemp) {
    double cumulative_time = 0.0;
    int i_heating = 0;
    double avg = 0.0;
    double intermediate_temperature = initial_temperature;
    for (;;) {
      i_heating = i_heating + 1;
      double cooling = coolin(s, ndens_atom, ndens_electron, intermediate_temperature) + cosmo_cool_rate;
      double thermal_rate = fmax(1e-50, fabs(cooling - heating));
      double thermal_timescale = internal_energy / fabs(thermal_rate);
      double dt_thermal = K_relative_denergy * thermal_timescale;
      double dt_ODE = fmin(dt_thermal, dt - cumulative_time);
      internal_energy = internal_energy + dt_ODE * (heating - cooling);
      avg = avg + F32(0.5f) * intermediate_temperature * dt_ODE;
      /* pressr2temper tped.f90:58-70: pressr/(k_B*(ndens+eldens)) */
      intermediate_temperature = internal_energy * K_gamma1 / (K_k_B * (ndens_atom + electrondens(ndens_atom, h_av)));
      avg = avg + F32(0.5f) * intermediate_temperature * dt_ODE;
      if (intermediate_temperature < K_minitemp) {
        internal_energy = (ndens_atom + electrondens(ndens_atom, h_av)) * K_k_B * K_minitemp; /* no /gamma1, :131 */
        intermediate_temperature = K_minitemp;
      }
      cumulative_time = cumulative_time + dt_ODE;
      if (cumulative_time >= dt || fabs(cumulative_time - dt) < F32(1e-6f) * dt) break;
      if (i_heating > 10000) break;
    }
    if (dt > 0.0) avg = avg / dt;
    else avg = initial_temperature;
    *average_temperature = avg;
    *final_temperature = internal_energy * K_gamma1 / (K_k_B * (ndens_atom + electrondens(ndens_atom, h)));
  }
}

/* evolve0D_global evolve_point.F90:305-406 with do_chemistry :410-555 (local=.false.) */
static void evolve0D_global(orc_state *s, double dt, size_t p, int *conv_flag) {
  double h[2], h_old[2], h_av[2];
  h[1] = fmax(K_epsilon, s->xh_intermed[p]);
  h_old[1] = fmax(K_epsilon, s->xh[p]);
  h_av[1] = fmax(K_epsilon, s->xh_av[p]);
  h[0] = F32(1.0f) - h[1];
  h_old[0] = F32(1.0f) - h_old[1];
  h_av[0] = F32(1.0f) - h_av[1];
  double ndens_p = (double)s->ndens[p];
  /* get_temperature_point temperature_module.F90:134-151: (current, average, intermed) */
  double Tstart_cur, Tstart_avg, Tstart_int;
  if (s->isothermal) {
    Tstart_cur = Tstart_avg = Tstart_int = s->temper_val;
  } else {
    Tstart_cur = (double)s->temperature_grid[3 * p];
    Tstart_avg = (double)s->temperature_grid[3 * p + 1];
    Tstart_int = (double)s->temperature_grid[3 * p + 2];
  }
  double phi_cell = s->phih_grid[p];
  double heat = s->isothermal ? 0.0 : s->phiheat_grid[p]; /* :364 */
  /* do_chemistry: temperature_end=temperature_start */
  double Tend_cur = Tstart_cur, Tend_avg = Tstart_avg, Tend_int = Tstart_int;
  float clumping = s->clumping;
  if (s->type_of_clumping == 3 || s->type_of_clumping == 4 || s->type_of_clumping == 5)
    clumping = s->clumping_grid[p]; /* clumping_point clumping_module.F90:106-118 */
  int nit = 0;
  for (;;) {
    nit = nit + 1;
    double temperature_previous_iteration = Tend_cur;
    double yh0_av_old = h_av[0];
    h[0] = h_old[0];
    h[1] = h_old[1];
    double de = electrondens(ndens_p, h_av);
    /* ini_rec_colion_factors(temperature_end%average) (:491) only sets module variables doric does not use */
    doric(dt, Tend_avg, de, ndens_p, h, h_av, phi_cell, clumping);
    de = electrondens(ndens_p, h_av);
    if (!s->isothermal) /* :519-526 */
      thermal(s, dt, Tstart_cur, &Tend_int, &Tend_avg, de, ndens_p, h, h_old, h_av, heat);
    /* the temperature term compares temperature_end%current, which thermal never updates (Appendix E.8) */
    if ((fabs((h_av[0] - yh0_av_old) / h_av[0]) < K_minimum_fractional_change ||
         h_av[0] < K_minimum_fraction_of_atoms) &&
        fabs((Tend_cur - temperature_previous_iteration) / Tend_cur) < K_minimum_fractional_change)
      break;
    if (nit > 400) break; /* :541-549 'Convergence failing (global)' */
  }
  /* set_temperature_point (:553, temperature_module.F90:155-169): intermed and average, as default real */
  if (!s->isothermal) {
    s->temperature_grid[3 * p + 2] = (float)Tend_int;
    s->temperature_grid[3 * p + 1] = (float)Tend_avg;
  }
  /* :378-391 */
  double yh1_av_old = fmax(K_epsilon, s->xh_av[p]);
  double yh0_av_old = F32(1.0f) - yh1_av_old;
  /* get_temperature_point again: the values just stored, rounded to default real */
  double Tnew_avg = s->isothermal ? s->temper_val : (double)s->temperature_grid[3 * p + 1];
  if ((fabs(h_av[0] - yh0_av_old) > K_minimum_fractional_change &&
       fabs((h_av[0] - yh0_av_old) / h_av[0]) > K_minimum_fractional_change &&
       h_av[0] > K_minimum_fraction_of_atoms) ||
      (fabs((Tstart_avg - Tnew_avg) / Tnew_avg) > 1.0e-1 && fabs(Tstart_avg - Tnew_avg) > 100.0))
    *conv_flag = *conv_flag + 1;
  s->xh_intermed[p] = h[1];
  s->xh_av[p] = h_av[1];
}

/* ------------------------------------------------------------------------------------------ */
/* photonstatistics.F90                                                                         */
/* ------------------------------------------------------------------------------------------ */
void orc_state_before(orc_state *s) { /* :104-132 */
  double h0 = 0.0, h1 = 0.0;
  for (size_t p = 0; p < s->ncell; ++p) {
    h0 = h0 + (double)s->ndens[p] * (1.0 - s->xh[p]);
    h1 = h1 + (double)s->ndens[p] * s->xh[p];
  }
  s->h0_before = h0 * s->vol;
  s->h1_before = h1 * s->vol;
}
static void state_after(orc_state *s, const double *xh_l) { /* :190-217 */
  double h0 = 0.0, h1 = 0.0;
  for (size_t p = 0; p < s->ncell; ++p) {
    h0 = h0 + (double)s->ndens[p] * (1.0 - xh_l[p]);
    h1 = h1 + (double)s->ndens[p] * xh_l[p];
  }
  s->h0_after = h0 * s->vol;
  s->h1_after = h1 * s->vol;
}
static void total_rates(orc_state *s, double dt, const double *xh_l) { /* :137-185 */
  double totrec = 0.0, totcoll = 0.0;
  const double T = s->temper_val;
  /* loop invariants evaluated once; every product below keeps the reference's left-to-right order */
  const double powT = pow(T / F32(1e4f), K_albpow), sqrtT = sqrt(T), expT = exp(-K_temph0() / T);
  for (size_t p = 0; p < s->ncell; ++p) {
    double yh[2];
    yh[0] = 1.0 - xh_l[p];
    yh[1] = xh_l[p];
    double ndens_p = (double)s->ndens[p];
    float clumping = s->clumping;
    if (s->type_of_clumping == 3 || s->type_of_clumping == 4 || s->type_of_clumping == 5)
      clumping = s->clumping_grid[p];
    if (s->isothermal) {
      totrec = totrec + ndens_p * yh[1] * electrondens(ndens_p, yh) * (double)clumping * K_bh00 * powT;
      totcoll = totcoll + ndens_p * yh[0] * electrondens(ndens_p, yh) * K_colh0() * sqrtT * expT;
    } else { /* temperature%average of the cell, photonstatistics.F90:166-176 */
      const double Ta = (double)s->temperature_grid[3 * p + 1];
      totrec = totrec + ndens_p * yh[1] * electrondens(ndens_p, yh) * (double)clumping * K_bh00 *
                            pow(Ta / F32(1e4f), K_albpow);
      totcoll = totcoll + ndens_p * yh[0] * electrondens(ndens_p, yh) * K_colh0() * sqrt(Ta) *
                              exp(-K_temph0() / Ta);
    }
  }
  s->totrec = totrec * s->vol * dt;
  s->totcollisions = totcoll * s->vol * dt;
}
static void fill_stats(const orc_state *s, double dt, orc_photon_stats *st) {
  st->h0_before = s->h0_before; st->h1_before = s->h1_before;
  st->h0_after = s->h0_after; st->h1_after = s->h1_after;
  st->totrec = s->totrec; st->totcollisions = s->totcollisions;
  st->dh0 = s->dh0; st->total_ion = s->total_ion;
  /* report_photonstatistics :254-281 */
  double m3 = (double)((float)s->mesh[0]);
  st->total_photon_loss = s->photon_loss * dt * m3 * (double)((float)s->mesh[1]) * (double)((float)s->mesh[2]);
  st->LLS_loss = s->LLS_loss;
  double sumflux = 0.0;
  for (int n = 0; n < s->NumSrc; ++n) sumflux = sumflux + s->NormFlux_stellar[n];
  st->totalsrc = sumflux * s->S_star * dt;
  st->photcons = (s->total_ion + s->LLS_loss - s->totcollisions) / st->totalsrc;
}
void orc_calculate_photon_statistics(orc_state *s, double dt, const double *xh_l, const double *xh_r,
                                     orc_photon_stats *st) { /* :82-99 */
  state_after(s, xh_l);
  total_rates(s, dt, xh_r);
  s->dh0 = (s->h0_before - s->h0_after); /* total_ionizations :222-228 */
  s->total_ion = s->totrec + s->dh0;
  if (st) fill_stats(s, dt, st);
}

/* ------------------------------------------------------------------------------------------ */
/* evolve.F90                                                                                   */
/* ------------------------------------------------------------------------------------------ */
/* global_pass :499-573 */
void orc_global_pass(orc_state *s, double dt, double photon_loss_all, orc_global_report *rep) {
  float m = (float)s->mesh[0] * (float)s->mesh[1] * (float)s->mesh[2];
  s->photon_loss = photon_loss_all / (double)m; /* :525 */
  double maxav = s->xh_av[0];
  for (size_t p = 1; p < s->ncell; ++p) if (s->xh_av[p] > maxav) maxav = s->xh_av[p];
  rep->min_avg_neutral = F32(1.0f) - maxav; /* :535 */
  int conv_flag = 0;
  /* :548-555, i fastest.  The cells are independent (conv_flag is an integer count), so with
   * orc_set_threads(n > 1) the loop is shared between host threads; the result is bit-identical. */
  const long long ncell = (long long)s->ncell;
#pragma omp parallel for num_threads(s->nthreads) reduction(+ : conv_flag) schedule(static)
  for (long long p = 0; p < ncell; ++p) evolve0D_global(s, dt, (size_t)p, &conv_flag);
  rep->conv_flag = conv_flag;
  double sum = 0.0;
  for (size_t p = 0; p < s->ncell; ++p) sum = sum + s->xh_intermed[p];
  rep->sum_xh_intermed = sum;
  orc_calculate_photon_statistics(s, dt, s->xh_intermed, s->xh_av, &rep->stats); /* :570 */
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* write_iteration_dump :285-324 as an in-memory record: niter | photon_loss_all | phih_grid | xh_av | xh_intermed */
static void write_iteration_dump(orc_state *s, int niter, double photon_loss_all) {
  const size_t b = s->ncell * sizeof(double);
  if (!s->dump_phih) {
    s->dump_phih = (double *)malloc(b);
    s->dump_xh_av = (double *)malloc(b);
    s->dump_xh_intermed = (double *)malloc(b);
  }
  s->dump_niter = niter;
  s->dump_photon_loss_all = photon_loss_all;
  memcpy(s->dump_phih, s->phih_grid, b);
  memcpy(s->dump_xh_av, s->xh_av, b);
  memcpy(s->dump_xh_intermed, s->xh_intermed, b);
  s->have_dump = 1;
}

/* The reference dumps when 15 minutes of wall clock have passed (:253-266); here the caller names the
 * iteration after whose pass_all_sources the dump is taken (0 = never). */
void orc_set_dump_iteration(orc_state *s, int niter) { s->dump_at_iter = niter; }

int orc_get_dump(const orc_state *s, int *niter, double *photon_loss_all, double *phih, double *xh_av,
                 double *xh_intermed) {
  if (!s->have_dump) return 1;
  const size_t b = s->ncell * sizeof(double);
  *niter = s->dump_niter;
  *photon_loss_all = s->dump_photon_loss_all;
  memcpy(phih, s->dump_phih, b);
  memcpy(xh_av, s->dump_xh_av, b);
  memcpy(xh_intermed, s->dump_xh_intermed, b);
  return 0;
}

/* evolve3D :83-281.  restart != 0: start_from_dump (:328-426) has put niter, photon_loss_all, phih_grid, xh_av
 * and xh_intermed in place (orc_evolve3D_restart), then global_pass runs before the loop (:154-158). */
static void evolve3D_impl(orc_state *s, double dt, int max_outer_iter, int restart, int niter0,
                          double photon_loss_all0, orc_step_report *rep) {
  memset(rep, 0, sizeof(*rep));
  orc_state_before(s); /* :136 */
  int niter = 0;
  int conv_flag = s->mesh[0] * s->mesh[1] * s->mesh[2];
  double prev_sum_xh1_int = (double)(2.0f * (float)s->mesh[0] * (float)s->mesh[1] * (float)s->mesh[2]);
  double prev_sum_xh0_int = prev_sum_xh1_int;
  double rel_change_sum_xh1 = 1.0, rel_change_sum_xh0 = 1.0;
  if (restart == 0) {
    memcpy(s->xh_av, s->xh, s->ncell * sizeof(double));       /* :140-147 */
    memcpy(s->xh_intermed, s->xh, s->ncell * sizeof(double));
  } else {
    /* the restart branch leaves prev_sum_xh*_int as they are: module variables without initialiser (:69-70),
     * zero in a freshly started run, which is when a restart happens */
    prev_sum_xh1_int = 0.0;
    prev_sum_xh0_int = 0.0;
    niter = niter0;
    orc_global_report gr0;
    orc_global_pass(s, dt, photon_loss_all0, &gr0); /* :157 */
    conv_flag = gr0.conv_flag;
  }
  /* :162-163 */
  int c1 = (int)(K_convergence_fraction * (double)s->mesh[0] * (double)s->mesh[1] * (double)s->mesh[2]);
  int c2 = (s->NumSrc - 1) / 3;
  int conv_criterion = c1 < c2 ? c1 : c2;
  rep->conv_criterion = conv_criterion;
  for (;;) {
    double sum_xh1_int = 0.0;
    for (size_t p = 0; p < s->ncell; ++p) sum_xh1_int = sum_xh1_int + s->xh_intermed[p]; /* :183 */
    double sum_xh0_int = (double)((float)(s->mesh[0] * s->mesh[1] * s->mesh[2])) - sum_xh1_int;
    if (sum_xh1_int > 0.0) rel_change_sum_xh1 = fabs(sum_xh1_int - prev_sum_xh1_int) / sum_xh1_int;
    else rel_change_sum_xh1 = 1.0;
    if (sum_xh0_int > 0.0) rel_change_sum_xh0 = fabs(sum_xh0_int - prev_sum_xh0_int) / sum_xh0_int;
    else rel_change_sum_xh0 = 1.0;
    if (niter < ORC_MAX_ITER) {
      rep->rel_change_sum_xh1[niter] = rel_change_sum_xh1;
      rep->rel_change_sum_xh0[niter] = rel_change_sum_xh0;
    }
    if (conv_flag < conv_criterion ||
        (rel_change_sum_xh1 < K_convergence_fraction && rel_change_sum_xh0 < K_convergence_fraction)) { /* :212-214 */
      memcpy(s->xh, s->xh_intermed, s->ncell * sizeof(double));
      if (!s->isothermal) /* set_final_temperature_point temperature_module.F90:173-183: current=intermed */
        for (size_t p = 0; p < s->ncell; ++p) s->temperature_grid[3 * p] = s->temperature_grid[3 * p + 2];
      rep->converged = 1;
      break;
    } else {
      if (niter > 100 || (max_outer_iter > 0 && niter >= max_outer_iter)) { /* :228-232 */
        rep->converged = 0;
        break;
      }
    }
    prev_sum_xh1_int = sum_xh1_int;
    prev_sum_xh0_int = sum_xh0_int;
    niter = niter + 1;
    orc_set_rates_to_zero(s);
    orc_pass_report pr;
    double t0 = now_s();
    orc_pass_all_sources(s, &pr);
    double t1 = now_s();
    if (s->dump_at_iter == niter) write_iteration_dump(s, niter, pr.photon_loss_all); /* :253-266 */
    orc_global_report gr;
    orc_global_pass(s, dt, pr.photon_loss_all, &gr);
    double t2 = now_s();
    rep->seconds_raytrace += t1 - t0;
    rep->seconds_global += t2 - t1;
    conv_flag = gr.conv_flag;
    if (niter < ORC_MAX_ITER) {
      rep->conv_flag[niter] = conv_flag;
      rep->photon_loss_all[niter] = pr.photon_loss_all;
      rep->sum_nbox_all[niter] = pr.sum_nbox_all;
      rep->updates[niter] = pr.updates;
      rep->iter_stats[niter] = gr.stats;
    }
    rep->total_updates += pr.updates;
  }
  rep->niter = niter;
  orc_calculate_photon_statistics(s, dt, s->xh, s->xh_av, &rep->final_stats); /* :277 */
  /* update_grandtotal_photonstatistics photonstatistics.F90:286-293 */
  double sumflux = 0.0;
  for (int n = 0; n < s->NumSrc; ++n) sumflux = sumflux + s->NormFlux_stellar[n];
  s->grtotal_src = s->grtotal_src + sumflux * s->S_star * dt;
  s->grtotal_ion = s->grtotal_ion + s->total_ion - s->totcollisions;
  rep->grtotal_ion = s->grtotal_ion;
  rep->grtotal_src = s->grtotal_src;
}

void orc_evolve3D(orc_state *s, double dt, int max_outer_iter, orc_step_report *rep) {
  evolve3D_impl(s, dt, max_outer_iter, 0, 0, 0.0, rep);
}

/* evolve3D(time,dt,restart/=0): the dump record is passed in instead of read from iterdump[12].bin */
void orc_evolve3D_restart(orc_state *s, double dt, int max_outer_iter, int niter, double photon_loss_all,
                          const double *phih, const double *xh_av, const double *xh_intermed,
                          orc_step_report *rep) {
  const size_t b = s->ncell * sizeof(double);
  memcpy(s->phih_grid, phih, b);
  memcpy(s->xh_av, xh_av, b);
  memcpy(s->xh_intermed, xh_intermed, b);
  evolve3D_impl(s, dt, max_outer_iter, 1, niter, photon_loss_all, rep);
}
