// Shared declarations of the B200-native C2-Ray hot path (internal; the public ABI is
// include/c2ray_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/c2ray_b200.h"

namespace c2b {

constexpr int kNumTau = C2B_NUMTAU;
constexpr int kTableLen = kNumTau + 1;

// ---- ray tracer ------------------------------------------------------------------------------
// One work item = one source.  A source is traced shell by shell (Chebyshev distance r from the
// source); the shell is stored as 6 faces (principal axis p, sign of the principal offset) of 4
// quadrants (signs of the two transverse offsets), each a square (r+1) x (r+1) patch of the plane
// |d_p| = r indexed by the transverse distances (a, b); the four quadrant values of (a,b) are
// contiguous.  A cell of plane r depends only on the four cells (a-1|a, b-1|b) of plane r-1 of the
// same quadrant (column_density.f90:108-171).
struct RtParams {
  int n[3];                 // mesh
  int lim[3][2];            // [axis][0: negative side L, 1: positive side R], evolve_source.F90:100-102
  int subboxsize;
  int plane_stride;         // S = max(lim)+1: a global plane buffer holds 6 faces of raytrace_face_doubles(S)
  int smem_plane_doubles;   // capacity of one shared-memory plane buffer (single-CTA kernel)
  int smem_plane_doubles_cl;  // same for a CTA of the cluster kernel
  const double* tau_cell;   // sigma_HI*dr(1)*max(1-max(xh_av,eps),eps)*ndens per cell
  double* phih;             // evolve_data.F90:40
  const double* tau_cell_t; // y-fastest twins used by the x-principal quadrants
  double* phih_t;
  const float* lls_grid;    // LLS.F90:81 (type_of_LLS == 2) or nullptr
  const double2* thick2;    // stellar_photo_thick_table as (value, forward difference) pairs
  const double* thin;       // stellar_photo_thin_table(0:NumTau,1)
  const double2* logtab;    // 128 x {1/c_j, A + B*log2(c_j)}
  // non-isothermal path (null when isothermal): stellar_heat_thick_table as pairs, ..thin.., phiheat_grid + twin
  const double2* heat2;
  const double* heat_thin;
  double* phiheat;
  double* phiheat_t;
  double tau_heat_limit;    // radiation_photoionrates.F90:333
  const int* srcpos;        // 3 x NumSrc, 1-based (sourceprops.F90:56)
  const double* normflux;   // NormFlux_stellar(1:NumSrc)
  const int* work;          // source indices (0-based) this rank traces, in order
  const int* nseg_cta;      // per shell radius: b-segments per column (single-CTA kernel / cluster kernel)
  const int* nseg_cl;
  const int* nseg_w;        // same for the per-warp kernel (32 threads walk the six faces)
  int warp_plane_doubles[2];   // capacities of the two shared-memory plane buffers (even / odd shells) of a face in the per-warp kernel
  int* ovf;                 // sources the per-warp kernel hands over to the single-CTA kernel after their first subbox
  unsigned int* ovf_count;  // ... and how many (written by the per-warp kernel; read by the single-CTA kernel when non-null)
  int nwork;
  unsigned int* ticket;     // dynamic work counter (plays do_grid_master, master_slave.F90:124-231)
  double* scratch;          // per-work-group global plane storage: [grid][2][6 faces][face_doubles]
  int* nbox_out;            // per source (global index)
  double* loss_out;         // per source
  double* coldens_dbg;      // optional full coldensh_out grid (debug/parity), or nullptr
  double S_star, dr[3], vol;
  double dr2[3];            // dr*dr per axis
  double tau_stop;          // sigma_HI*max_coldensh (evolve_point.F90:95,201)
  double vol_cell;          // dr(1)*dr(2)*dr(3), vol_ph of the source cell (:153)
  int use_lls, type_lls;
  int cubic_cells;          // dr(1)==dr(2)==dr(3)
  int use_twins;            // the x-principal faces of the CTA / cluster kernels work on the y-fastest twins (else: strided, on the x-fastest grids)
  double tau_lls;           // sigma_HI*coldensh_LLS
  double rmax_lls2;
  double sigma_HI, inv_sigma, inv_sigma_dr0, fourpi_over_sigma;
  double max_coldensh, tau_photo_limit, loss_fraction;
  double sqrt2, sqrt3;
  double logB;              // log10(2)/dlogtau
  double logc[5];           // logB/ln2 * {1,-1/2,1/3,-1/4,1/5}
};

void launch_raytrace(const RtParams& p, int grid, cudaStream_t stream);
struct RtLaunchInfo {
  int smem_plane_doubles, smem_plane_doubles_cl;
  int grid_cta;    // resident CTAs of the single-CTA kernel
  int clusters;    // resident clusters of the cluster kernel
  int cluster_size;
  int grid_max;    // CTAs the scratch must be sized for
  int warp_warps;  // warps (= concurrent sources) per CTA of the per-warp kernel; 0: not usable for this mesh / subboxsize
  int warp_plane_doubles[2];
  int grid_warp;   // CTAs of the per-warp kernel (one per SM)
};
// sets the kernels' shared-memory attributes and queries the resident grid sizes
// min_lim = the smallest half-box limit (evolve_source.F90:100-102): the per-warp kernel needs subboxsize < min_lim
int raytrace_configure(int max_radius, bool heat_tables, int subboxsize, int min_lim, RtLaunchInfo* info);
int launch_raytrace_cluster(const RtParams& p, int nclusters, cudaStream_t stream);
void launch_raytrace_warp(const RtParams& p, int grid, int warps, cudaStream_t stream);
void raytrace_nseg_tables(int max_radius, std::vector<int>& cta, std::vector<int>& cl, std::vector<int>& warp);
void launch_taucell(const float* ndens, const double* xh_av, double* tau_cell, size_t n, double sigma_dr0,
                    double eps, cudaStream_t stream);
void launch_pair_table(const double* tab, double2* out, cudaStream_t stream);
size_t raytrace_scratch_doubles_per_cta(int plane_stride);
// doubles of one face (4 quadrants of S x S cells) of a global plane buffer, with slack for the reads of
// zero-weight neighbours just outside the plane
__host__ __device__ inline size_t raytrace_face_doubles(int plane_stride) {
  return (size_t)4 * plane_stride * plane_stride + (size_t)8 * plane_stride + 32;
}

// ---- per-cell chemistry + fused statistics ------------------------------------------------------
constexpr int kNumStat = 8;
// slots of the per-block partials
enum StatSlot { kSumXh = 0, kH0 = 1, kH1 = 2, kRec = 3, kColl = 4, kMaxXhAv = 5, kConv = 6, kSpare = 7 };

struct ChemParams {
  size_t ncell;
  const float* ndens;
  const double* xh;        // ionfractions_module.F90:22 (start of step)
  double* xh_av;           // in: previous iterate, out: new time average
  double* xh_intermed;     // out
  const double* phih;
  const float* clumping_grid;  // type 3/4/5 or nullptr
  float clumping;              // scalar (default real, clumping_module.F90:17)
  double dt, inv_dt;       // inv_dt = 1/dt (0 when dt == 0: statistics-only launches)
  double bh00_powT;        // bh00*(T/1e4)**albpow is formed as (clumping*bh00)*powT, doric.f90:74
  double bh00, powT;
  double acolh0;           // colh0*sqrt(T)*exp(-temph0/T), doric.f90:77
  double colh0, sqrtT, expT;  // the same three factors, multiplied per cell in total_rates order
  double abu_c, epsilon, minimum_fractional_change, minimum_fraction_of_atoms;
  double* partials;        // [nblocks][kNumStat]
  double* tau_cell;        // out (chemistry kernel): opacity grid for the next ray trace, or nullptr
  double sigma_dr0;
  // non-isothermal path (all null / unused when isothermal): temperature_grid as three float arrays
  const float* T_cur;      // temperature_grid%current (start of the step)
  float* T_avg;            // %average: in = previous iterate, out = new time average
  float* T_int;            // %intermed: out
  const double* phiheat;   // evolve_data.F90:42
  const double* cie_cool;  // 61 entries, cooling.f90:26-27
  double cool_mintemp, cool_dtemp;
  double k_B, gamma1, minitemp, relative_denergy;
  double cosmo_cool_factor;   // 2/(1+zred)*dzdt (cosmology.F90:198-225), 0 when not cosmological
  double temph0, albpow;
  double temper_val;       // isothermal temperature (unused otherwise)
};

void launch_chemistry(const ChemParams& p, int nblocks, cudaStream_t stream);
// statistics only: sum(x_l), h0/h1 from (ndens,x_l); totrec/totcoll from (ndens,x_r) (x_r may be null)
void launch_stats(const ChemParams& p, const double* x_l, const double* x_r, int nblocks,
                  cudaStream_t stream);
// temperature_grid (current, average, intermed per cell, default real) <-> three float arrays
void launch_unpack_temperature(const float* aos, float* cur, float* avg, float* inter, size_t n, cudaStream_t stream);
void launch_pack_temperature(const float* cur, const float* avg, const float* inter, float* aos, size_t n, cudaStream_t stream);
void launch_fill_f32(float* a, float v, size_t n, cudaStream_t stream);
void launch_clumping_from_density(const float* ndens, float* clump, size_t n, double p1, double p2, double p3,
                                  double avg_dens, cudaStream_t stream);
void launch_finalize_partials(const double* partials, int nblocks, double* out /*kNumStat*/,
                              cudaStream_t stream);
void launch_scale_density(float* ndens, size_t n, double zfactor3, cudaStream_t stream);
void launch_to_f32(const double* in, float* out, size_t n, cudaStream_t stream);
void launch_to_yfast(const double* in, double* out, const int n[3], cudaStream_t stream);
void launch_add_from_yfast(double* acc, const double* twin, const int n[3], cudaStream_t stream);
int chemistry_blocks();

// ---- rate tables (rad_ini) -----------------------------------------------------------------------
struct SedParams {
  double T_eff, S_star, freq_min, freq_max, hplanck, k_B, two_pi_over_c_square, R_solar, pi;
  double pl_index_cross_section;
  double minlogtau, dlogtau;
};
int build_blackbody_tables(const SedParams& sp, double* d_thick, double* d_thin, double* d_heat_thick,
                           double* d_heat_thin, cudaStream_t stream, double* S_star_unscaled_out);

double measure_dfma_rate(cudaStream_t stream);

}  // namespace c2b
