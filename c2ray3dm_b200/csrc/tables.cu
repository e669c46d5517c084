// rad_ini for a black-body SED: builds stellar_photo_thick_table / stellar_photo_thin_table
// (radiation_tables.F90:95-236, 361-452, 524-543) with the Romberg weights of romberg.f90:22-90
// and the SED normalisation of radiation_sed_parameters.F90:82-283.
//
// The frequency-side vectors (129 entries) are prepared on the host; the 2001 x 129 integrand
// evaluations and the weighted sums run on the device, one thread per optical-depth entry, adding
// the frequency points in the same order as vector_romberg (romberg.f90:158-187).
#include <cmath>
#include <vector>

#include "c2b_common.cuh"

namespace c2b {
namespace {

constexpr int kNumFreq = 128;  // radiation_sizes.f90:13

struct FreqSide {
  double hnu[kNumFreq + 1];     // hplanck*(frequency-ion_freq_HI), radiation_tables.F90:482-487
  double sed[kNumFreq + 1];     // BB_SED(i_freq)
  double cs[kNumFreq + 1];      // cross_section_freq_dependence
  double wgt[kNumFreq + 1];     // romw(:,7)
  double delta_freq;
};

__constant__ FreqSide c_freq;

// heat_thick / heat_thin may be null (isothermal)
__global__ void photo_table_kernel(double* thick, double* thin, double* heat_thick, double* heat_thin, double minlogtau,
                                   double dlogtau) {
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it > kNumTau) return;
  // tau(0)=0 ; tau(i)=10.0**(minlogtau+dlogtau*real(i-1))  (radiation_tables.F90:148-155)
  const double tau = (it == 0) ? 0.0 : pow(10.0, minlogtau + dlogtau * (double)(it - 1));
  double a_thick = 0.0, a_thin = 0.0, h_thick = 0.0, h_thin = 0.0;
  for (int i = 0; i <= kNumFreq; ++i) {
    const double x = tau * c_freq.cs[i];
    double f_thick = 0.0, f_thin = 0.0;
    if (x < 700.0) {  // :386-399
      const double e = exp(-x);
      f_thick = c_freq.sed[i] * e;
      f_thin = c_freq.sed[i] * c_freq.cs[i] * e;
    }
    a_thick = a_thick + f_thick * c_freq.delta_freq * c_freq.wgt[i];
    a_thin = a_thin + f_thin * c_freq.delta_freq * c_freq.wgt[i];
    // fill_heating_integrands_HI + make_heat_tables_HI (radiation_tables.F90:471-509, 547-565)
    h_thick = h_thick + (c_freq.hnu[i] * f_thick) * c_freq.delta_freq * c_freq.wgt[i];
    h_thin = h_thin + (c_freq.hnu[i] * f_thin) * c_freq.delta_freq * c_freq.wgt[i];
  }
  thick[it] = a_thick;
  thin[it] = a_thin;
  if (heat_thick) heat_thick[it] = h_thick;
  if (heat_thin) heat_thin[it] = h_thin;
}

// Romberg weights for 2**pmax intervals: superposition of the Richardson-extrapolated
// trapezoid rules, as romberg_initialisation builds them (romberg.f90:22-90).  The extrapolation
// coefficient -1/(4**k-1) is a default-real expression there, hence the float arithmetic.
void romberg_weights(int pmax, std::vector<double>& w) {
  const int n = 1 << pmax;
  std::vector<double> a(pmax + 1, 0.0), b(pmax + 1, 0.0);
  for (int k = 1; k <= pmax; ++k) {
    const float four_k = std::ldexp(1.0f, 2 * k);
    b[k] = (double)(-1.0f / (four_k - 1.0f));
    a[k] = -b[k] * (double)four_k;
  }
  w.assign(n + 1, 0.0);
  std::vector<std::vector<double>> s(pmax + 1, std::vector<double>(pmax + 1, 0.0));
  for (int level = 0; level <= pmax; ++level) {
    // coefficient with which the trapezoid rule of 2**level intervals enters the final estimate
    for (auto& row : s) std::fill(row.begin(), row.end(), 0.0);
    s[level][0] = 1.0;
    for (int j = 1; j <= pmax; ++j)
      for (int i = pmax; i >= j; --i) s[i][j] = a[j] * s[i][j - 1] + b[j] * s[i - 1][j - 1];
    const int step = 1 << (pmax - level);
    for (int j = 0; j <= (1 << level); ++j) w[(size_t)step * j] = s[pmax][pmax] * (double)step + w[(size_t)step * j];
  }
  w[0] *= 0.5;
  w[n] *= 0.5;
}

}  // namespace

int build_blackbody_tables(const SedParams& sp, double* d_thick, double* d_thin, double* d_heat_thick,
                           double* d_heat_thin, cudaStream_t stream, double* S_star_unscaled_out) {
  std::vector<double> w;
  romberg_weights(7, w);
  const double h_over_kT = sp.hplanck / (sp.k_B * sp.T_eff);
  const double step = (sp.freq_max - sp.freq_min) / (double)(float)kNumFreq;
  // spec_diag / integrate_sed("B","S"): photon rate of a black body of radius R_solar
  double integral = 0.0;
  for (int i = 0; i <= kNumFreq; ++i) {
    const double f = sp.freq_min + step * (double)(float)i;
    double g;
    if (f * h_over_kT <= 709.0) g = sp.two_pi_over_c_square * f * f / (std::exp(f * h_over_kT) - 1.0);
    else g = sp.two_pi_over_c_square * f * f / std::exp((f * h_over_kT) / 2.0) / std::exp((f * h_over_kT) / 2.0);
    integral = integral + g * step * w[i] * 1.0;
  }
  const double S_unscaled = 4.0 * sp.pi * sp.R_solar * sp.R_solar * integral;
  const double S_scaling = sp.S_star / S_unscaled;
  const double R_star = std::sqrt(S_scaling) * sp.R_solar;
  const double R_star2 = R_star * R_star;
  if (S_star_unscaled_out) *S_star_unscaled_out = S_unscaled;

  FreqSide fs;
  fs.delta_freq = step;
  for (int i = 0; i <= kNumFreq; ++i) {
    const double f = sp.freq_min + step * (double)(float)i;
    fs.cs[i] = std::pow(f / sp.freq_min, -sp.pl_index_cross_section);
    fs.sed[i] = (f * h_over_kT < 700.0)
                    ? 4.0 * sp.pi * R_star2 * sp.two_pi_over_c_square * f * f / (std::exp(f * h_over_kT) - 1.0)
                    : 0.0;
    fs.wgt[i] = w[i];
    fs.hnu[i] = sp.hplanck * (f - sp.freq_min);   // freq_min = ion_freq_HI (radiation_sizes.f90:61)
  }
  cudaError_t e = cudaMemcpyToSymbolAsync(c_freq, &fs, sizeof(fs), 0, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return (int)e;
  photo_table_kernel<<<(kTableLen + 127) / 128, 128, 0, stream>>>(d_thick, d_thin, d_heat_thick, d_heat_thin, sp.minlogtau,
                                                                  sp.dlogtau);
  e = cudaStreamSynchronize(stream);  // fs lives on this stack frame
  return (int)e;
}

}  // namespace c2b
