// C ABI of the B200-native C2-Ray hot path: handle, module-state marshalling, the evolve3D outer
// loop (evolve.F90:83-281), pass_all_sources (evolve.F90:444-495 + master_slave.F90:74-96),
// global_pass (evolve.F90:499-573), the photon statistics bookkeeping (photonstatistics.F90) and
// the rank reduction (evolve.F90:577-616) over NCCL.  See include/c2ray_b200.h.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "c2b_common.cuh"

using namespace c2b;

namespace {

thread_local std::string g_create_error;

struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      err = std::string("cannot load NCCL: ") + dlerror();
      return false;
    }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GetErrorString) {
      err = "NCCL library lacks a required symbol";
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;

}  // namespace

struct c2b_handle {
  c2b_config cfg;
  size_t ncell = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_step[2] = {nullptr, nullptr};   // brackets of c2b_evolve3d
  std::string error;
  // grids
  float* d_ndens = nullptr;
  double *d_xh = nullptr, *d_xh_av = nullptr, *d_xh_intermed = nullptr, *d_phih = nullptr;
  float *d_clump = nullptr, *d_lls = nullptr, *d_f32tmp = nullptr;
  double *d_thick = nullptr, *d_thin = nullptr, *d_taucell = nullptr;
  double *d_taucell_t = nullptr, *d_phih_t = nullptr;   // y-fastest twins for the x-principal quadrants
  double* d_xh_saved = nullptr;                         // c2b_save_xh_dev
  // non-isothermal path
  double *d_phiheat = nullptr, *d_phiheat_t = nullptr;  // evolve_data.F90:42 and its y-fastest twin
  double *d_heat_thick = nullptr, *d_heat_thin = nullptr, *d_cie_cool = nullptr;
  double2* d_heat2 = nullptr;
  float *d_Tcur = nullptr, *d_Tavg = nullptr, *d_Tint = nullptr, *d_Taos = nullptr;   // temperature_grid
  double cool_mintemp = 0.0, cool_dtemp = 1.0, zred = 0.0;
  bool have_heat_tables = false, have_cooling = false;
  bool taucell_t_dirty = true;
  double2 *d_thick2 = nullptr, *d_logtab = nullptr;
  bool taucell_dirty = true;   // xh_av / ndens / dr changed since tau_cell was last formed
  RtLaunchInfo rt;             // shared-memory plane capacities and resident grid sizes
  int* d_work2 = nullptr;      // work list of the cluster kernel
  int* d_work3 = nullptr;      // work list of the per-warp kernel (traces that ended after one subbox last time)
  int* d_ovf = nullptr;        // sources the per-warp kernel hands over to the single-CTA kernel
  int warp_min_sources = 0;    // the per-warp kernel is used when at least this many sources qualify
  unsigned int* h_ovf = nullptr;   // pinned: number of handed-over sources of the last pass
  // the three work lists of the last pass: they stay on the device while no source's predicted length changes
  bool routes_valid = false;
  double est_updates = 0.0;     // predicted updates of the CTA + cluster kernels for the cached routes
  int n_small = 0, n_large = 0, n_tiny = 0;
  std::vector<int64_t> updates_of_nbox;   // cells of the final subbox for nbox = 0, 1, 2, ...
  long long route_counts[4] = {0, 0, 0, 0};
  int *d_nseg_cta = nullptr, *d_nseg_cl = nullptr, *d_nseg_w = nullptr;
  std::vector<int> nbox_pred;  // per source: nbox of the previous trace (routing + longest-first order)
  int cluster_min_nbox = 3;    // sources predicted to need >= this many subboxes may go to the cluster kernel
  int cluster_max_sources = 0; // ... but only while there are too few of them to fill the GPU one CTA each
  bool have_tables = false, have_density = false, have_xh = false, have_geometry = false;
  // sources
  int NumSrc = 0, nwork = 0;
  int* d_srcpos = nullptr;
  double* d_normflux = nullptr;
  int* d_work = nullptr;
  int* d_nbox = nullptr;
  double* d_loss = nullptr;
  int* h_nbox = nullptr;      // pinned
  double* h_loss = nullptr;   // pinned
  // multi-rank load balance: which rank traces which source in the next pass (the static round-robin of
  // master_slave.F90:85 until the trace lengths and the ranks' measured speeds are known)
  std::vector<int> owner;          // per source
  std::vector<int> morton_all;     // all sources in Z-order of their mesh position
  std::vector<double> rank_speed;  // updates per ms of every rank's last ray-trace pass (as dealt with)
  bool balance = false;            // nranks > 1 and not disabled
  long long redeals = 0;
  int* h_nbox_all = nullptr;       // pinned: nbox of every source after the all-reduce
  std::vector<int> work;      // 0-based source indices of this rank
  std::vector<int> work_morton;  // the same indices in Z-order of their mesh position (L2 locality of concurrent traces)
  std::vector<int> srcpos;
  double sum_normflux = 0.0, S_star = 0.0;
  // ray-trace launch resources
  unsigned int* d_ticket = nullptr;
  double* d_scratch = nullptr;
  int rt_grid = 0, plane_stride = 0;
  int lim[3][2];
  // reductions
  double* d_partials = nullptr;
  double* d_stats = nullptr;
  double* h_stats = nullptr;  // pinned, kNumStat
  double* d_small = nullptr;  // 4 doubles for the packed scalar all-reduce
  double* h_small = nullptr;  // pinned
  int chem_blocks = 0;
  // module state
  double dr[3] = {0, 0, 0}, vol = 0.0;
  float clumping = 1.0f;
  double coldensh_LLS = 0.0, R_max_LLS = 0.0, temper_val = 1.0e4;
  double h0_before = 0, h1_before = 0, h0_after = 0, h1_after = 0, totrec = 0, totcoll = 0, dh0 = 0,
         total_ion = 0;
  double photon_loss = 0.0, LLS_loss = 0.0, grtotal_ion = 0.0, grtotal_src = 0.0;
  double sum_xh_intermed = 0.0;
  // restart state
  int iter_niter = 0;
  double iter_photon_loss_all = 0.0;
  bool have_iter_state = false;
  // NCCL
  ncclComm_t comm = nullptr;
  long long launches = 0;
};

#define C2B_CHECK_H(h)            \
  do {                            \
    if (!(h)) return 100;         \
  } while (0)

#define CU(h, call)                                                                     \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      char buf__[512];                                                                  \
      snprintf(buf__, sizeof(buf__), "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, \
               cudaGetErrorString(e__));                                                \
      (h)->error = buf__;                                                               \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

#define NC(h, call)                                                                     \
  do {                                                                                  \
    ncclResult_t r__ = (call);                                                          \
    if (r__ != ncclSuccess) {                                                           \
      char buf__[512];                                                                  \
      snprintf(buf__, sizeof(buf__), "%s:%d: %s failed: %s", __FILE__, __LINE__, #call, \
               g_nccl.GetErrorString(r__));                                             \
      (h)->error = buf__;                                                               \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)

static int fail(c2b_handle* h, const char* msg) {
  h->error = msg;
  return 3;
}

static int bind_device(c2b_handle* h) {
  CU(h, cudaSetDevice(h->cfg.device));
  return 0;
}

extern "C" {

int c2b_default_config(c2b_config* c) {
  if (!c) return 100;
  memset(c, 0, sizeof(*c));
  c->mesh[0] = c->mesh[1] = c->mesh[2] = 300;              // sizes.f90:33
  c->device = 0;
  c->rank = 0;
  c->nranks = 1;
  c->isothermal = 1;                                       // c2ray_parameters.f90:28
  c->type_of_clumping = 1;                                 // :75
  c->use_LLS = 1;                                          // :80
  c->type_of_LLS = 1;                                      // :87
  c->subboxsize = 5;                                       // :54
  c->max_subbox = 1000;                                    // :61
  c->max_outer_iter = 100;                                 // evolve.F90:228
  c->epsilon = 1e-14;                                      // :31
  c->convergence_fraction = (double)1.0e-4f;               // :25 (default-real literal)
  c->minimum_fractional_change = (double)1.0e-3f;          // :34
  c->minimum_fraction_of_atoms = (double)1.0e-8f;          // :40
  c->loss_fraction = 1e-2;                                 // :67
  c->max_coldensh = (double)2e19f;                         // evolve_point.F90:95
  c->tau_photo_limit = (double)1.0e-7f;                    // radiation_photoionrates.F90:244
  c->minlogtau = -20.0;                                    // radiation_tables.F90:45
  c->dlogtau = (4.0 - (-20.0)) / (double)(float)C2B_NUMTAU;  // :47
  c->sigma_HI = 1.0 * (double)6.30e-18f;                   // cgsphotoconstants.f90:24
  c->pi = (double)3.141592654f;                            // mathconstants.f90:21
  c->sqrt2 = (double)sqrtf(2.0f);                          // column_density.f90:53
  c->sqrt3 = (double)sqrtf(3.0f);                          // :52
  c->bh00 = 2.59e-13;                                      // cgsconstants.f90:66
  c->albpow = -0.7;                                        // :64
  const double eth0 = (double)13.598f;                     // :76
  const double ev2k = (double)(1.0f / 8.617e-05f);         // :39
  c->temph0 = eth0 * ev2k;                                 // :80
  c->colh0 = (double)1.3e-8f * (double)0.83f * (double)1.0f / (eth0 * eth0);  // :86
  c->abu_c = (double)7.1e-7f;                              // abundances.f90:26
  c->k_B = 1.381e-16;                                      // cgsconstants.f90:34
  c->gamma1 = 5.0 / 3.0 - 1.0;                             // atomic.f90:23-25
  c->minitemp = (double)1.0f;                              // c2ray_parameters.f90:108
  c->relative_denergy = (double)0.1f;                      // :110
  c->tau_heat_limit = (double)1.0e-4f;                     // radiation_photoionrates.F90:333
  const double Mpc = (double)1e6f * (double)3.086e18f;     // cgsastroconstants.f90:29-31
  c->H0 = (double)0.7f * (double)100.0f * (double)1e5f / Mpc;   // cosmoparms.f90:28,41
  c->Omega0 = (double)0.27f;                               // :30
  c->cosmological = 1;                                     // c2ray_parameters.f90:105
  return 0;
}

int c2b_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char* c2b_last_error(const c2b_handle* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int c2b_create(const c2b_config* cfg, c2b_handle** out) {
  if (!cfg || !out) {
    g_create_error = "c2b_create: null argument";
    return 100;
  }
  *out = nullptr;
  for (int d = 0; d < 3; ++d)
    if (cfg->mesh[d] < 4 || cfg->mesh[d] > 4096) {
      g_create_error = "c2b_create: mesh must be within [4, 4096] per axis";
      return 101;
    }
  if ((double)cfg->mesh[0] * (double)cfg->mesh[1] * (double)cfg->mesh[2] >= 2147483648.0) {
    g_create_error = "c2b_create: mesh(1)*mesh(2)*mesh(3) must be below 2^31 (the reference's default-integer cell counts, evolve.F90:148,162)";
    return 101;
  }
  if (cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks) {
    g_create_error = "c2b_create: bad rank/nranks";
    return 103;
  }
  if (cfg->use_LLS && (cfg->type_of_LLS < 1 || cfg->type_of_LLS > 3)) {
    g_create_error = "c2b_create: type_of_LLS must be 1, 2 or 3";
    return 104;
  }
  if (cfg->type_of_clumping < 1 || cfg->type_of_clumping > 5 || cfg->subboxsize < 1 || cfg->max_subbox < 1) {
    g_create_error = "c2b_create: bad type_of_clumping / subboxsize / max_subbox";
    return 105;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev < 1) {
    g_create_error = std::string("c2b_create: no CUDA device (") + cudaGetErrorString(e) +
                     "); this library has no CPU fallback";
    return 110;
  }
  if (cfg->device < 0 || cfg->device >= ndev) {
    g_create_error = "c2b_create: device ordinal out of range";
    return 111;
  }
  c2b_handle* h = new c2b_handle();
  h->cfg = *cfg;
  h->ncell = (size_t)cfg->mesh[0] * cfg->mesh[1] * cfg->mesh[2];
  auto bail = [&](const char* what, cudaError_t ce) {
    g_create_error = std::string("c2b_create: ") + what + ": " + cudaGetErrorString(ce);
    c2b_destroy(h);
    return 112;
  };
  if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess)
    return bail("cudaStreamCreate", e);
  for (auto& ev : h->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
  for (auto& ev : h->ev_step)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
  const size_t n = h->ncell;
  if ((e = cudaMalloc(&h->d_ndens, n * sizeof(float))) != cudaSuccess) return bail("cudaMalloc ndens", e);
  if ((e = cudaMalloc(&h->d_xh, n * sizeof(double))) != cudaSuccess) return bail("cudaMalloc xh", e);
  if ((e = cudaMalloc(&h->d_xh_av, n * sizeof(double))) != cudaSuccess) return bail("cudaMalloc xh_av", e);
  if ((e = cudaMalloc(&h->d_xh_intermed, n * sizeof(double))) != cudaSuccess) return bail("cudaMalloc xh_intermed", e);
  if ((e = cudaMalloc(&h->d_phih, n * sizeof(double))) != cudaSuccess) return bail("cudaMalloc phih", e);
  if ((e = cudaMemsetAsync(h->d_phih, 0, n * sizeof(double), h->stream)) != cudaSuccess) return bail("memset", e);
  if ((e = cudaMalloc(&h->d_thick, kTableLen * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&h->d_thin, kTableLen * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
  // [0] single-CTA kernel, [1] cluster kernel, [2] per-warp kernel, [3] number of handed-over sources
  if ((e = cudaMalloc(&h->d_ticket, 4 * sizeof(unsigned int))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&h->d_taucell, n * sizeof(double))) != cudaSuccess) return bail("cudaMalloc tau_cell", e);
  if ((e = cudaMalloc(&h->d_taucell_t, n * sizeof(double))) != cudaSuccess) return bail("cudaMalloc tau_cell_t", e);
  if ((e = cudaMalloc(&h->d_phih_t, n * sizeof(double))) != cudaSuccess) return bail("cudaMalloc phih_t", e);
  if ((e = cudaMalloc(&h->d_thick2, kTableLen * sizeof(double2))) != cudaSuccess) return bail("cudaMalloc", e);
  if (!cfg->isothermal) {
    if ((e = cudaMalloc(&h->d_phiheat, n * sizeof(double))) != cudaSuccess) return bail("cudaMalloc phiheat", e);
    if ((e = cudaMalloc(&h->d_phiheat_t, n * sizeof(double))) != cudaSuccess) return bail("cudaMalloc phiheat_t", e);
    if ((e = cudaMemsetAsync(h->d_phiheat, 0, n * sizeof(double), h->stream)) != cudaSuccess) return bail("memset", e);
    if ((e = cudaMalloc(&h->d_heat_thick, kTableLen * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&h->d_heat_thin, kTableLen * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&h->d_heat2, kTableLen * sizeof(double2))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&h->d_cie_cool, 64 * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&h->d_Tcur, n * sizeof(float))) != cudaSuccess) return bail("cudaMalloc T", e);
    if ((e = cudaMalloc(&h->d_Tavg, n * sizeof(float))) != cudaSuccess) return bail("cudaMalloc T", e);
    if ((e = cudaMalloc(&h->d_Tint, n * sizeof(float))) != cudaSuccess) return bail("cudaMalloc T", e);
    if ((e = cudaMalloc(&h->d_Taos, 3 * n * sizeof(float))) != cudaSuccess) return bail("cudaMalloc T", e);
    // temperature_array_init (temperature_module.F90:44-67) with the default temper_val
    launch_fill_f32(h->d_Tcur, (float)h->temper_val, n, h->stream);
    launch_fill_f32(h->d_Tavg, (float)h->temper_val, n, h->stream);
    launch_fill_f32(h->d_Tint, (float)h->temper_val, n, h->stream);
  }
  if ((e = cudaMalloc(&h->d_logtab, 128 * sizeof(double2))) != cudaSuccess) return bail("cudaMalloc", e);
  {
    // log2 table of the table-coordinate evaluation (raytrace.cu: table_coord)
    const double B = std::log10(2.0) / cfg->dlogtau, A = 1.0 - cfg->minlogtau / cfg->dlogtau;
    double2 lt[128];
    for (int j = 0; j < 128; ++j) {
      const double cj = 1.0 + (j + 0.5) / 128.0;
      lt[j].x = 1.0 / cj;
      lt[j].y = A + B * std::log2(cj);
    }
    if ((e = cudaMemcpy(h->d_logtab, lt, sizeof(lt), cudaMemcpyHostToDevice)) != cudaSuccess) return bail("memcpy", e);
  }
  // evolve_source.F90:100-102 (periodic_bc)
  int smax = 0;
  for (int d = 0; d < 3; ++d) {
    h->lim[d][1] = std::min(cfg->max_subbox, cfg->mesh[d] / 2 - 1 + cfg->mesh[d] % 2);
    h->lim[d][0] = std::min(cfg->max_subbox, cfg->mesh[d] / 2);
    smax = std::max(smax, std::max(h->lim[d][0], h->lim[d][1]));
  }
  h->plane_stride = smax + 1;
  int smin = smax;
  for (int d = 0; d < 3; ++d) smin = std::min(smin, std::min(h->lim[d][0], h->lim[d][1]));
  if (raytrace_configure(smax, !cfg->isothermal, cfg->subboxsize, smin, &h->rt)) return bail("raytrace_configure", cudaGetLastError());
  h->warp_min_sources = 2 * h->rt.grid_warp;
  if (const char* env = getenv("C2B_WARP_MIN_SOURCES")) h->warp_min_sources = atoi(env);
  h->rt_grid = h->rt.grid_max;
  h->cluster_max_sources = 4 * h->rt.clusters;
  if (const char* env = getenv("C2B_CLUSTER_MIN_NBOX")) h->cluster_min_nbox = atoi(env);
  if (const char* env = getenv("C2B_CLUSTER_MAX_SOURCES")) h->cluster_max_sources = atoi(env);
  {
    std::vector<int> t_cta, t_cl, t_w;
    raytrace_nseg_tables(smax, t_cta, t_cl, t_w);
    if ((e = cudaMalloc(&h->d_nseg_w, t_w.size() * sizeof(int))) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemcpy(h->d_nseg_w, t_w.data(), t_w.size() * sizeof(int), cudaMemcpyHostToDevice);
    if ((e = cudaMalloc(&h->d_nseg_cta, t_cta.size() * sizeof(int))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&h->d_nseg_cl, t_cl.size() * sizeof(int))) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemcpy(h->d_nseg_cta, t_cta.data(), t_cta.size() * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_nseg_cl, t_cl.data(), t_cl.size() * sizeof(int), cudaMemcpyHostToDevice);
  }
  const size_t scratch = raytrace_scratch_doubles_per_cta(h->plane_stride) * (size_t)h->rt_grid;
  if ((e = cudaMalloc(&h->d_scratch, scratch * sizeof(double))) != cudaSuccess) return bail("cudaMalloc scratch", e);
  // the ray tracer reads zero-weight neighbours outside the planes: they must be finite (raytrace.cu)
  if ((e = cudaMemsetAsync(h->d_scratch, 0, scratch * sizeof(double), h->stream)) != cudaSuccess) return bail("memset scratch", e);
  h->chem_blocks = chemistry_blocks();
  if ((e = cudaMalloc(&h->d_partials, (size_t)h->chem_blocks * kNumStat * sizeof(double))) != cudaSuccess)
    return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&h->d_stats, kNumStat * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
  const size_t nsmall = 4 + 2 * (size_t)cfg->nranks;   // {loss, nbox, updates, -} + {ms, updates} of every rank
  if ((e = cudaMalloc(&h->d_small, nsmall * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMallocHost(&h->h_stats, kNumStat * sizeof(double))) != cudaSuccess) return bail("cudaMallocHost", e);
  if ((e = cudaMallocHost(&h->h_small, nsmall * sizeof(double))) != cudaSuccess) return bail("cudaMallocHost", e);
  if ((e = cudaMallocHost(&h->h_ovf, sizeof(unsigned int))) != cudaSuccess) return bail("cudaMallocHost", e);
  *h->h_ovf = 0u;
  if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return bail("sync", e);
  *out = h;
  return 0;
}

void c2b_destroy(c2b_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_ndens); cudaFree(h->d_xh); cudaFree(h->d_xh_av); cudaFree(h->d_xh_intermed);
  cudaFree(h->d_phih); cudaFree(h->d_clump); cudaFree(h->d_lls); cudaFree(h->d_f32tmp);
  cudaFree(h->d_thick); cudaFree(h->d_thin); cudaFree(h->d_taucell); cudaFree(h->d_taucell_t); cudaFree(h->d_phih_t); cudaFree(h->d_thick2); cudaFree(h->d_logtab); cudaFree(h->d_nseg_cta); cudaFree(h->d_nseg_cl); cudaFree(h->d_nseg_w); cudaFree(h->d_srcpos); cudaFree(h->d_normflux);
  cudaFree(h->d_work); cudaFree(h->d_work2); cudaFree(h->d_work3); cudaFree(h->d_ovf); cudaFree(h->d_nbox); cudaFree(h->d_loss); cudaFree(h->d_ticket);
  cudaFree(h->d_xh_saved);
  cudaFree(h->d_phiheat); cudaFree(h->d_phiheat_t); cudaFree(h->d_heat_thick); cudaFree(h->d_heat_thin); cudaFree(h->d_heat2);
  cudaFree(h->d_cie_cool); cudaFree(h->d_Tcur); cudaFree(h->d_Tavg); cudaFree(h->d_Tint); cudaFree(h->d_Taos);
  cudaFree(h->d_scratch); cudaFree(h->d_partials); cudaFree(h->d_stats); cudaFree(h->d_small);
  if (h->h_nbox) cudaFreeHost(h->h_nbox);
  if (h->h_loss) cudaFreeHost(h->h_loss);
  if (h->h_nbox_all) cudaFreeHost(h->h_nbox_all);
  if (h->h_stats) cudaFreeHost(h->h_stats);
  if (h->h_small) cudaFreeHost(h->h_small);
  if (h->h_ovf) cudaFreeHost(h->h_ovf);
  for (auto& ev : h->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : h->ev_step)
    if (ev) cudaEventDestroy(ev);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

// ---- multi-GPU ---------------------------------------------------------------------------------
int c2b_get_unique_id(void* id) {
  if (!id) return 100;
  std::string err;
  if (!g_nccl.load(err)) {
    g_create_error = err;
    return 2;
  }
  static_assert(sizeof(ncclUniqueId) <= C2B_UNIQUE_ID_BYTES, "unique id size");
  ncclUniqueId uid;
  if (g_nccl.GetUniqueId(&uid) != ncclSuccess) {
    g_create_error = "ncclGetUniqueId failed";
    return 2;
  }
  memset(id, 0, C2B_UNIQUE_ID_BYTES);
  memcpy(id, &uid, sizeof(uid));
  return 0;
}

int c2b_comm_init(c2b_handle* h, const void* id) {
  C2B_CHECK_H(h);
  if (!id) return fail(h, "c2b_comm_init: null id");
  if (h->cfg.nranks == 1) return 0;
  if (!g_nccl.load(h->error)) return 2;
  if (bind_device(h)) return 1;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  NC(h, g_nccl.CommInitRank(&h->comm, h->cfg.nranks, uid, h->cfg.rank));
  return 0;
}

// ---- inputs -------------------------------------------------------------------------------------
int c2b_set_tables(c2b_handle* h, const double* thick, const double* thin, int32_t n) {
  C2B_CHECK_H(h);
  if (!thick || !thin) return fail(h, "c2b_set_tables: null table");
  if (n != kTableLen) return fail(h, "c2b_set_tables: n must be NumTau+1 = 2001");
  if (bind_device(h)) return 1;
  CU(h, cudaMemcpyAsync(h->d_thick, thick, kTableLen * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_thin, thin, kTableLen * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  launch_pair_table(h->d_thick, h->d_thick2, h->stream);
  h->launches += 1;
  CU(h, cudaStreamSynchronize(h->stream));
  h->have_tables = true;
  return 0;
}

int c2b_rad_ini_blackbody(c2b_handle* h, double T_eff, double S_star, double freq_min, double freq_max,
                          double hplanck, double k_B, double two_pi_over_c_square, double R_solar,
                          double pl_index_cross_section, double* thick_out, double* thin_out) {
  C2B_CHECK_H(h);
  if (bind_device(h)) return 1;
  SedParams sp;
  sp.T_eff = T_eff; sp.S_star = S_star; sp.freq_min = freq_min; sp.freq_max = freq_max;
  sp.hplanck = hplanck; sp.k_B = k_B; sp.two_pi_over_c_square = two_pi_over_c_square;
  sp.R_solar = R_solar; sp.pi = h->cfg.pi; sp.pl_index_cross_section = pl_index_cross_section;
  sp.minlogtau = h->cfg.minlogtau; sp.dlogtau = h->cfg.dlogtau;
  int rc = build_blackbody_tables(sp, h->d_thick, h->d_thin, h->d_heat_thick, h->d_heat_thin, h->stream, nullptr);
  h->launches += 1;
  if (rc) return fail(h, "c2b_rad_ini_blackbody: table kernel failed");
  launch_pair_table(h->d_thick, h->d_thick2, h->stream);
  h->launches += 1;
  if (h->d_heat_thick) {
    launch_pair_table(h->d_heat_thick, h->d_heat2, h->stream);
    h->launches += 1;
    h->have_heat_tables = true;
  }
  CU(h, cudaStreamSynchronize(h->stream));
  h->have_tables = true;
  if (thick_out) CU(h, cudaMemcpy(thick_out, h->d_thick, kTableLen * sizeof(double), cudaMemcpyDeviceToHost));
  if (thin_out) CU(h, cudaMemcpy(thin_out, h->d_thin, kTableLen * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int c2b_set_density(c2b_handle* h, const float* ndens) {
  C2B_CHECK_H(h);
  if (!ndens) return fail(h, "c2b_set_density: null pointer");
  if (bind_device(h)) return 1;
  CU(h, cudaMemcpyAsync(h->d_ndens, ndens, h->ncell * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  h->have_density = true;
  h->taucell_dirty = true;
  return 0;
}

int c2b_set_geometry(c2b_handle* h, const double dr[3], double vol) {
  C2B_CHECK_H(h);
  if (!dr) return fail(h, "c2b_set_geometry: null pointer");
  if (!(dr[0] > 0) || !(dr[1] > 0) || !(dr[2] > 0) || !(vol > 0)) return fail(h, "c2b_set_geometry: non-positive dr/vol");
  h->dr[0] = dr[0]; h->dr[1] = dr[1]; h->dr[2] = dr[2];
  h->vol = vol;
  h->have_geometry = true;
  h->taucell_dirty = true;
  return 0;
}

int c2b_cosmo_evol(c2b_handle* h, double zfactor) {
  C2B_CHECK_H(h);
  if (!(zfactor > 0)) return fail(h, "c2b_cosmo_evol: zfactor must be positive");
  if (!h->have_density || !h->have_geometry) return fail(h, "c2b_cosmo_evol: density/geometry not set");
  if (bind_device(h)) return 1;
  const double zfactor3 = zfactor * zfactor * zfactor;  // cosmology.F90:174
  for (int d = 0; d < 3; ++d) h->dr[d] = h->dr[d] * zfactor;
  h->vol = h->vol * zfactor3;
  launch_scale_density(h->d_ndens, h->ncell, zfactor3, h->stream);
  h->launches += 1;
  h->taucell_dirty = true;
  CU(h, cudaGetLastError());
  return 0;
}

int c2b_set_clumping_scalar(c2b_handle* h, float clumping) {
  C2B_CHECK_H(h);
  h->clumping = clumping;
  return 0;
}

static int upload_f32_grid(c2b_handle* h, float** dst, const float* src) {
  if (bind_device(h)) return 1;
  if (!*dst) CU(h, cudaMalloc(dst, h->ncell * sizeof(float)));
  CU(h, cudaMemcpyAsync(*dst, src, h->ncell * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int c2b_set_clumping_grid(c2b_handle* h, const float* g) {
  C2B_CHECK_H(h);
  if (!g) return fail(h, "c2b_set_clumping_grid: null pointer");
  if (h->cfg.type_of_clumping < 3) return fail(h, "c2b_set_clumping_grid: type_of_clumping is 1 or 2 (scalar)");
  return upload_f32_grid(h, &h->d_clump, g);
}

int c2b_set_clumping_from_density(c2b_handle* h, double p1, double p2, double p3, double avg_dens) {
  C2B_CHECK_H(h);
  if (h->cfg.type_of_clumping < 3) return fail(h, "c2b_set_clumping_from_density: type_of_clumping is 1 or 2 (scalar)");
  if (!h->have_density) return fail(h, "c2b_set_clumping_from_density: density not set (c2b_set_density)");
  if (!(avg_dens > 0.0)) return fail(h, "c2b_set_clumping_from_density: avg_dens must be positive");
  if (bind_device(h)) return 1;
  if (!h->d_clump) CU(h, cudaMalloc(&h->d_clump, h->ncell * sizeof(float)));
  launch_clumping_from_density(h->d_ndens, h->d_clump, h->ncell, p1, p2, p3, avg_dens, h->stream);
  h->launches += 1;
  CU(h, cudaGetLastError());
  return 0;
}

int c2b_get_clumping_grid(c2b_handle* h, float* g) {
  C2B_CHECK_H(h);
  if (!g) return fail(h, "c2b_get_clumping_grid: null pointer");
  if (!h->d_clump) return fail(h, "c2b_get_clumping_grid: no clumping grid on the device");
  if (bind_device(h)) return 1;
  CU(h, cudaMemcpyAsync(g, h->d_clump, h->ncell * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int c2b_set_lls_scalar(c2b_handle* h, double v) {
  C2B_CHECK_H(h);
  h->coldensh_LLS = v;
  return 0;
}
int c2b_set_lls_grid(c2b_handle* h, const float* g) {
  C2B_CHECK_H(h);
  if (!g) return fail(h, "c2b_set_lls_grid: null pointer");
  if (h->cfg.type_of_LLS != 2) return fail(h, "c2b_set_lls_grid: type_of_LLS is not 2");
  return upload_f32_grid(h, &h->d_lls, g);
}
int c2b_set_lls_rmax(c2b_handle* h, double v) {
  C2B_CHECK_H(h);
  h->R_max_LLS = v;
  return 0;
}
int c2b_set_temperature(c2b_handle* h, double t) {
  C2B_CHECK_H(h);
  if (!(t > 0)) return fail(h, "c2b_set_temperature: temperature must be positive");
  h->temper_val = t;
  if (h->d_Tcur) {   // temperature_array_init, temperature_module.F90:44-67
    if (bind_device(h)) return 1;
    launch_fill_f32(h->d_Tcur, (float)t, h->ncell, h->stream);
    launch_fill_f32(h->d_Tavg, (float)t, h->ncell, h->stream);
    launch_fill_f32(h->d_Tint, (float)t, h->ncell, h->stream);
    h->launches += 3;
    CU(h, cudaGetLastError());
  }
  return 0;
}

// Deals sources to ranks for one pass: cost[s] = predicted updates of source s, speed[r] = relative speed of rank r.
// Longest trace first, each to the rank that is furthest below its share speed[r]/sum(speed) of the total cost
// (ties: the lower rank).  Deterministic: every rank computes the same assignment from the same all-reduced inputs.
int c2b_deal_sources(int32_t nsrc, const int64_t* cost, int32_t nranks, const double* speed, int32_t* owner) {
  if (nsrc < 0 || nranks < 1 || (nsrc > 0 && (!cost || !owner))) return 1;
  double stot = 0.0, ctot = 0.0;
  for (int r = 0; r < nranks; ++r) stot += (speed && speed[r] > 0.0) ? speed[r] : 1.0;
  for (int s = 0; s < nsrc; ++s) ctot += (double)std::max<int64_t>(cost[s], 1);
  std::vector<double> deficit((size_t)nranks);
  for (int r = 0; r < nranks; ++r) deficit[(size_t)r] = ctot * ((speed && speed[r] > 0.0) ? speed[r] : 1.0) / stot;
  std::vector<int> order((size_t)nsrc);
  for (int s = 0; s < nsrc; ++s) order[(size_t)s] = s;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  for (int s : order) {
    int best = 0;
    for (int r = 1; r < nranks; ++r)
      if (deficit[(size_t)r] > deficit[(size_t)best]) best = r;
    owner[s] = best;
    deficit[(size_t)best] -= (double)std::max<int64_t>(cost[s], 1);
  }
  return 0;
}

// this rank's sources from h->owner, in source order (work) and in Z-order (work_morton)
static void rebuild_work(c2b_handle* h) {
  h->work.clear();
  h->work_morton.clear();
  for (int s = 0; s < h->NumSrc; ++s)
    if (h->owner[(size_t)s] == h->cfg.rank) h->work.push_back(s);
  for (int s : h->morton_all)
    if (h->owner[(size_t)s] == h->cfg.rank) h->work_morton.push_back(s);
  h->nwork = (int)h->work.size();
  h->routes_valid = false;
}

int c2b_set_sources(c2b_handle* h, int32_t NumSrc, const int32_t* srcpos, const double* nf, double S_star) {
  C2B_CHECK_H(h);
  if (NumSrc < 0) return fail(h, "c2b_set_sources: negative NumSrc");
  if (NumSrc > 0 && (!srcpos || !nf)) return fail(h, "c2b_set_sources: null pointer");
  for (int s = 0; s < NumSrc; ++s)
    for (int d = 0; d < 3; ++d)
      if (srcpos[3 * s + d] < 1 || srcpos[3 * s + d] > h->cfg.mesh[d])
        return fail(h, "c2b_set_sources: source position outside the mesh (positions are 1-based)");
  if (bind_device(h)) return 1;
  cudaFree(h->d_srcpos); cudaFree(h->d_normflux); cudaFree(h->d_work); cudaFree(h->d_work2); cudaFree(h->d_work3); cudaFree(h->d_ovf); cudaFree(h->d_nbox); cudaFree(h->d_loss);
  h->d_work2 = nullptr; h->d_work3 = nullptr; h->d_ovf = nullptr;
  h->d_srcpos = nullptr; h->d_normflux = nullptr; h->d_work = nullptr; h->d_nbox = nullptr; h->d_loss = nullptr;
  if (h->h_nbox) cudaFreeHost(h->h_nbox);
  if (h->h_loss) cudaFreeHost(h->h_loss);
  if (h->h_nbox_all) cudaFreeHost(h->h_nbox_all);
  h->h_nbox = nullptr; h->h_loss = nullptr; h->h_nbox_all = nullptr;
  // the same source list again (a host that re-sends its state every step): keep the per-source trace lengths
  // the work queue is ordered by
  const bool same_sources = NumSrc == h->NumSrc && NumSrc > 0 && (int)h->nbox_pred.size() == NumSrc &&
                            std::equal(srcpos, srcpos + 3 * (size_t)NumSrc, h->srcpos.begin());
  std::vector<int> kept_pred, kept_owner;
  if (same_sources) kept_pred = h->nbox_pred;
  if (same_sources && (int)h->owner.size() == NumSrc) kept_owner = h->owner;
  h->NumSrc = NumSrc;
  h->S_star = S_star;
  h->srcpos.assign(srcpos, srcpos + 3 * (size_t)NumSrc);
  // do ns1=1+rank,NumSrc,npr (master_slave.F90:85)
  h->owner.assign((size_t)NumSrc, 0);
  for (int s = 0; s < NumSrc; ++s) h->owner[(size_t)s] = s % h->cfg.nranks;
  if (!kept_owner.empty()) h->owner = kept_owner;
  h->balance = h->cfg.nranks > 1 && !(getenv("C2B_NO_BALANCE") && atoi(getenv("C2B_NO_BALANCE")));
  if (h->rank_speed.size() != (size_t)h->cfg.nranks) h->rank_speed.assign((size_t)h->cfg.nranks, 1.0);
  h->nbox_pred.assign((size_t)NumSrc, 0);
  if (same_sources) h->nbox_pred = kept_pred;
  {
    // Z-order (Morton) key of the source cell: consecutive work items are spatial neighbours, so the traces
    // in flight at the same time share grid lines in L2
    auto spread = [](unsigned long long v) {
      v &= 0x1fffffULL;
      v = (v | (v << 32)) & 0x1f00000000ffffULL;
      v = (v | (v << 16)) & 0x1f0000ff0000ffULL;
      v = (v | (v << 8)) & 0x100f00f00f00f00fULL;
      v = (v | (v << 4)) & 0x10c30c30c30c30c3ULL;
      v = (v | (v << 2)) & 0x1249249249249249ULL;
      return v;
    };
    std::vector<std::pair<unsigned long long, int>> keyed;
    keyed.reserve((size_t)NumSrc);
    for (int w = 0; w < NumSrc; ++w) {
      const unsigned long long k = spread((unsigned)(srcpos[3 * w] - 1) >> 3) | (spread((unsigned)(srcpos[3 * w + 1] - 1) >> 3) << 1) |
                                   (spread((unsigned)(srcpos[3 * w + 2] - 1) >> 3) << 2);
      keyed.emplace_back(k, w);
    }
    std::stable_sort(keyed.begin(), keyed.end());
    h->morton_all.clear();
    for (auto& kw : keyed) h->morton_all.push_back(kw.second);
  }
  rebuild_work(h);
  h->sum_normflux = 0.0;
  for (int s = 0; s < NumSrc; ++s) h->sum_normflux = h->sum_normflux + nf[s];  // sum(NormFlux_stellar(1:NumSrc))
  if (NumSrc == 0) return 0;
  const size_t ns = (size_t)NumSrc;
  CU(h, cudaMalloc(&h->d_srcpos, 3 * ns * sizeof(int)));
  CU(h, cudaMalloc(&h->d_normflux, ns * sizeof(double)));
  // (any rank may be dealt any number of the sources)
  CU(h, cudaMalloc(&h->d_work, ns * sizeof(int)));
  CU(h, cudaMalloc(&h->d_work2, ns * sizeof(int)));
  CU(h, cudaMalloc(&h->d_work3, ns * sizeof(int)));
  CU(h, cudaMalloc(&h->d_ovf, ns * sizeof(int)));
  CU(h, cudaMalloc(&h->d_nbox, ns * sizeof(int)));
  CU(h, cudaMalloc(&h->d_loss, ns * sizeof(double)));
  CU(h, cudaMallocHost(&h->h_nbox, ns * sizeof(int)));
  CU(h, cudaMallocHost(&h->h_loss, ns * sizeof(double)));
  CU(h, cudaMallocHost(&h->h_nbox_all, ns * sizeof(int)));
  CU(h, cudaMemcpyAsync(h->d_srcpos, srcpos, 3 * ns * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_normflux, nf, ns * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (!h->work.empty())
    CU(h, cudaMemcpyAsync(h->d_work, h->work.data(), h->work.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemsetAsync(h->d_nbox, 0, ns * sizeof(int), h->stream));
  CU(h, cudaMemsetAsync(h->d_loss, 0, ns * sizeof(double), h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  memset(h->h_nbox, 0, ns * sizeof(int));
  memset(h->h_nbox_all, 0, ns * sizeof(int));
  memset(h->h_loss, 0, ns * sizeof(double));
  return 0;
}

int c2b_set_xh(c2b_handle* h, const double* xh) {
  C2B_CHECK_H(h);
  if (!xh) return fail(h, "c2b_set_xh: null pointer");
  if (bind_device(h)) return 1;
  CU(h, cudaMemcpyAsync(h->d_xh, xh, h->ncell * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  h->have_xh = true;
  return 0;
}

// ---- non-isothermal inputs ------------------------------------------------------------------------
static int need_thermal(c2b_handle* h, const char* who) {
  if (h->cfg.isothermal) {
    h->error = std::string(who) + ": the handle was created with isothermal=.true.";
    return 3;
  }
  return 0;
}

int c2b_set_heat_tables(c2b_handle* h, const double* thick, const double* thin, int32_t n) {
  C2B_CHECK_H(h);
  if (int rc = need_thermal(h, "c2b_set_heat_tables")) return rc;
  if (!thick || !thin) return fail(h, "c2b_set_heat_tables: null table");
  if (n != kTableLen) return fail(h, "c2b_set_heat_tables: n must be NumTau+1 = 2001");
  if (bind_device(h)) return 1;
  CU(h, cudaMemcpyAsync(h->d_heat_thick, thick, kTableLen * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_heat_thin, thin, kTableLen * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  launch_pair_table(h->d_heat_thick, h->d_heat2, h->stream);
  h->launches += 1;
  CU(h, cudaStreamSynchronize(h->stream));
  h->have_heat_tables = true;
  return 0;
}

int c2b_get_heat_tables(c2b_handle* h, double* thick, double* thin) {
  C2B_CHECK_H(h);
  if (int rc = need_thermal(h, "c2b_get_heat_tables")) return rc;
  if (!h->have_heat_tables) return fail(h, "c2b_get_heat_tables: heat tables not set");
  if (!thick || !thin) return fail(h, "c2b_get_heat_tables: null pointer");
  if (bind_device(h)) return 1;
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, cudaMemcpy(thick, h->d_heat_thick, kTableLen * sizeof(double), cudaMemcpyDeviceToHost));
  CU(h, cudaMemcpy(thin, h->d_heat_thin, kTableLen * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int c2b_set_cooling_table(c2b_handle* h, const double* log10_temp, const double* log10_cool, int32_t n) {
  C2B_CHECK_H(h);
  if (int rc = need_thermal(h, "c2b_set_cooling_table")) return rc;
  if (!log10_temp || !log10_cool) return fail(h, "c2b_set_cooling_table: null pointer");
  if (n != 61) return fail(h, "c2b_set_cooling_table: the CIE table has 61 rows (cooling.f90:26)");
  if (bind_device(h)) return 1;
  double cie[64] = {0};
  for (int i = 0; i < 61; ++i) cie[i] = std::pow(10.0, log10_cool[i]);   // cooling.f90:84-86
  h->cool_mintemp = log10_temp[0];                                       // :79-80
  h->cool_dtemp = log10_temp[1] - log10_temp[0];
  if (!(h->cool_dtemp > 0)) return fail(h, "c2b_set_cooling_table: temperatures must increase");
  CU(h, cudaMemcpy(h->d_cie_cool, cie, sizeof(cie), cudaMemcpyHostToDevice));
  h->have_cooling = true;
  return 0;
}

int c2b_set_redshift(c2b_handle* h, double zred) {
  C2B_CHECK_H(h);
  if (!(zred > -1.0)) return fail(h, "c2b_set_redshift: zred must exceed -1");
  h->zred = zred;
  return 0;
}

int c2b_set_temperature_grid(c2b_handle* h, const float* tg) {
  C2B_CHECK_H(h);
  if (int rc = need_thermal(h, "c2b_set_temperature_grid")) return rc;
  if (!tg) return fail(h, "c2b_set_temperature_grid: null pointer");
  if (bind_device(h)) return 1;
  CU(h, cudaMemcpyAsync(h->d_Taos, tg, 3 * h->ncell * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  launch_unpack_temperature(h->d_Taos, h->d_Tcur, h->d_Tavg, h->d_Tint, h->ncell, h->stream);
  h->launches += 1;
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int c2b_get_temperature_grid(c2b_handle* h, float* tg) {
  C2B_CHECK_H(h);
  if (int rc = need_thermal(h, "c2b_get_temperature_grid")) return rc;
  if (!tg) return fail(h, "c2b_get_temperature_grid: null pointer");
  if (bind_device(h)) return 1;
  launch_pack_temperature(h->d_Tcur, h->d_Tavg, h->d_Tint, h->d_Taos, h->ncell, h->stream);
  h->launches += 1;
  CU(h, cudaMemcpyAsync(tg, h->d_Taos, 3 * h->ncell * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

// ---- internals of the hot path ------------------------------------------------------------------
static void fill_chem(c2b_handle* h, double dt, ChemParams& cp) {
  const c2b_config& c = h->cfg;
  cp.ncell = h->ncell;
  cp.ndens = h->d_ndens;
  cp.xh = h->d_xh;
  cp.xh_av = h->d_xh_av;
  cp.xh_intermed = h->d_xh_intermed;
  cp.phih = h->d_phih;
  cp.clumping_grid = (c.type_of_clumping >= 3) ? h->d_clump : nullptr;
  cp.clumping = h->clumping;
  cp.dt = dt;
  cp.inv_dt = dt > 0.0 ? 1.0 / dt : 0.0;
  const double T = h->temper_val;
  cp.bh00 = c.bh00;
  cp.powT = std::pow(T / (double)1e4f, c.albpow);   // (temp0/1e4)**albpow, doric.f90:74
  cp.bh00_powT = c.bh00 * cp.powT;
  cp.sqrtT = std::sqrt(T);
  cp.expT = std::exp(-c.temph0 / T);
  cp.colh0 = c.colh0;
  cp.acolh0 = c.colh0 * cp.sqrtT * cp.expT;         // doric.f90:77
  cp.abu_c = c.abu_c;
  cp.epsilon = c.epsilon;
  cp.minimum_fractional_change = c.minimum_fractional_change;
  cp.minimum_fraction_of_atoms = c.minimum_fraction_of_atoms;
  cp.partials = h->d_partials;
  cp.tau_cell = h->d_taucell;
  cp.sigma_dr0 = c.sigma_HI * h->dr[0];
  cp.temph0 = c.temph0;
  cp.albpow = c.albpow;
  cp.temper_val = T;
  cp.T_cur = nullptr; cp.T_avg = nullptr; cp.T_int = nullptr; cp.phiheat = nullptr; cp.cie_cool = nullptr;
  cp.cool_mintemp = cp.cool_dtemp = cp.k_B = cp.gamma1 = cp.minitemp = cp.relative_denergy = cp.cosmo_cool_factor = 0.0;
  if (!c.isothermal) {
    cp.T_cur = h->d_Tcur; cp.T_avg = h->d_Tavg; cp.T_int = h->d_Tint;
    cp.phiheat = h->d_phiheat;
    cp.cie_cool = h->d_cie_cool;
    cp.cool_mintemp = h->cool_mintemp; cp.cool_dtemp = h->cool_dtemp;
    cp.k_B = c.k_B; cp.gamma1 = c.gamma1; cp.minitemp = c.minitemp; cp.relative_denergy = c.relative_denergy;
    if (c.cosmological) {   // cosmo_cool = e_int*2.0/(1.0+zred)*dzdt, cosmology.F90:198-225
      const double z1 = 1.0 + h->zred;
      const double dzdt = c.H0 * z1 * std::sqrt(c.Omega0 * (z1 * z1 * z1) + 1.0 - c.Omega0);
      cp.cosmo_cool_factor = 2.0 / z1 * dzdt;
    }
  }
}

static int fetch_stats(c2b_handle* h) {
  launch_finalize_partials(h->d_partials, h->chem_blocks, h->d_stats, h->stream);
  h->launches += 1;
  CU(h, cudaMemcpyAsync(h->h_stats, h->d_stats, kNumStat * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  CU(h, cudaGetLastError());
  return 0;
}

static int check_ready(c2b_handle* h) {
  if (!h->have_tables) return fail(h, "photo-ionization tables not set (c2b_set_tables / c2b_rad_ini_blackbody)");
  if (!h->have_density) return fail(h, "density not set (c2b_set_density)");
  if (!h->have_geometry) return fail(h, "geometry not set (c2b_set_geometry)");
  if (!h->have_xh) return fail(h, "ionization fractions not set (c2b_set_xh)");
  if (h->cfg.type_of_clumping >= 3 && !h->d_clump) return fail(h, "clumping grid not set");
  if (h->cfg.use_LLS && h->cfg.type_of_LLS == 2 && !h->d_lls) return fail(h, "LLS grid not set");
  if (h->cfg.nranks > 1 && !h->comm) return fail(h, "nranks > 1 but c2b_comm_init was not called");
  if (!h->cfg.isothermal) {
    if (!h->have_heat_tables) return fail(h, "heating tables not set (c2b_set_heat_tables / c2b_rad_ini_blackbody)");
    if (!h->have_cooling) return fail(h, "cooling table not set (c2b_set_cooling_table)");
  }
  return 0;
}

static void fill_photon_stats(const c2b_handle* h, double dt, c2b_photon_stats* st) {
  st->h0_before = h->h0_before; st->h1_before = h->h1_before;
  st->h0_after = h->h0_after; st->h1_after = h->h1_after;
  st->totrec = h->totrec; st->totcollisions = h->totcoll;
  st->dh0 = h->dh0; st->total_ion = h->total_ion;
  // report_photonstatistics, photonstatistics.F90:254-281
  st->total_photon_loss = h->photon_loss * dt * (double)(float)h->cfg.mesh[0] * (double)(float)h->cfg.mesh[1] *
                          (double)(float)h->cfg.mesh[2];
  st->LLS_loss = h->LLS_loss;
  st->totalsrc = h->sum_normflux * h->S_star * dt;
  st->photcons = (h->total_ion + h->LLS_loss - h->totcoll) / st->totalsrc;
}

// after a statistics reduction over (x_l, x_r): state_after + total_rates + total_ionizations
static void absorb_after(c2b_handle* h, double dt) {
  h->h0_after = h->h_stats[kH0] * h->vol;
  h->h1_after = h->h_stats[kH1] * h->vol;
  h->totrec = h->h_stats[kRec] * h->vol * dt;
  h->totcoll = h->h_stats[kColl] * h->vol * dt;
  h->dh0 = (h->h0_before - h->h0_after);
  h->total_ion = h->totrec + h->dh0;
}

// traces the sources of d_work (one CTA each), of d_work_cl (one cluster of 6 CTAs each) and of d_work_w (one warp
// each for the first subbox; those that need more are handed over to the single-CTA kernel on the device)
// novf > 0: follow-up call after a pass in which only the per-warp kernel ran and handed `novf` sources over (they are
// in d_ovf, their number in d_ticket[3]): the single-CTA kernel traces just those.
// The launches are bracketed by the events ev[e0], ev[e0+1]; the caller reads the time after its own synchronisation.
// est_updates: predicted updates of the CTA / cluster kernels in this call (decides whether the twin grids pay).
static int trace_sources(c2b_handle* h, const int* d_work, int nwork, const int* d_work_cl, int nwork_cl,
                         const int* d_work_w, int nwork_w, int novf, double* coldens_dbg, int e0,
                         double est_updates = 1e300) {
  const c2b_config& c = h->cfg;
  RtParams rp;
  memset(&rp, 0, sizeof(rp));
  for (int d = 0; d < 3; ++d) {
    rp.n[d] = c.mesh[d];
    rp.lim[d][0] = h->lim[d][0];
    rp.lim[d][1] = h->lim[d][1];
    rp.dr[d] = h->dr[d];
  }
  for (int d = 0; d < 3; ++d) rp.dr2[d] = h->dr[d] * h->dr[d];
  rp.tau_stop = c.max_coldensh * c.sigma_HI;
  rp.vol_cell = h->dr[0] * h->dr[1] * h->dr[2];
  rp.subboxsize = c.subboxsize;
  rp.plane_stride = h->plane_stride;
  rp.smem_plane_doubles = h->rt.smem_plane_doubles;
  rp.smem_plane_doubles_cl = h->rt.smem_plane_doubles_cl;
  rp.tau_cell = h->d_taucell;
  rp.tau_cell_t = h->d_taucell_t;
  rp.phih_t = h->d_phih_t;
  rp.phih = h->d_phih;
  rp.lls_grid = h->d_lls;
  rp.thick2 = h->d_thick2;
  rp.thin = h->d_thin;
  rp.logtab = h->d_logtab;
  rp.heat2 = h->d_heat2;
  rp.heat_thin = h->d_heat_thin;
  rp.phiheat = h->d_phiheat;
  rp.phiheat_t = h->d_phiheat_t;
  rp.tau_heat_limit = c.tau_heat_limit;
  rp.srcpos = h->d_srcpos;
  rp.normflux = h->d_normflux;
  rp.work = d_work;
  rp.nseg_cta = h->d_nseg_cta;
  rp.nseg_cl = h->d_nseg_cl;
  rp.nseg_w = h->d_nseg_w;
  rp.warp_plane_doubles[0] = h->rt.warp_plane_doubles[0];
  rp.warp_plane_doubles[1] = h->rt.warp_plane_doubles[1];
  rp.ovf = h->d_ovf;
  rp.ovf_count = nullptr;
  rp.nwork = nwork;
  rp.ticket = h->d_ticket;
  rp.scratch = h->d_scratch;
  rp.nbox_out = h->d_nbox;
  rp.loss_out = h->d_loss;
  rp.coldens_dbg = coldens_dbg;
  rp.S_star = h->S_star;
  rp.vol = h->vol;
  rp.use_lls = c.use_LLS;
  rp.cubic_cells = (h->dr[0] == h->dr[1] && h->dr[1] == h->dr[2]) ? 1 : 0;
  rp.type_lls = c.type_of_LLS;
  rp.tau_lls = (c.use_LLS && c.type_of_LLS == 1) ? c.sigma_HI * h->coldensh_LLS : 0.0;  // no LLS == a zero LLS column
  rp.rmax_lls2 = h->R_max_LLS * h->R_max_LLS;
  rp.sigma_HI = c.sigma_HI;
  rp.inv_sigma = 1.0 / c.sigma_HI;
  rp.inv_sigma_dr0 = 1.0 / (c.sigma_HI * h->dr[0]);
  rp.fourpi_over_sigma = 4.0 * c.pi / c.sigma_HI;
  rp.max_coldensh = c.max_coldensh;
  rp.tau_photo_limit = c.tau_photo_limit;
  rp.loss_fraction = c.loss_fraction;
  rp.sqrt2 = c.sqrt2;
  rp.sqrt3 = c.sqrt3;
  rp.logB = std::log10(2.0) / c.dlogtau;
  {
    const double k = rp.logB / std::log(2.0);
    rp.logc[0] = k; rp.logc[1] = -k / 2.0; rp.logc[2] = k / 3.0; rp.logc[3] = -k / 4.0; rp.logc[4] = k / 5.0;
  }
  if (h->taucell_dirty) {
    // opacity grid from the current xh_av (evolve_point.F90:137-145)
    launch_taucell(h->d_ndens, h->d_xh_av, h->d_taucell, h->ncell, c.sigma_HI * h->dr[0], c.epsilon, h->stream);
    h->launches += 1;
    h->taucell_dirty = false;
    h->taucell_t_dirty = true;
  }
  // the single-CTA and the cluster kernel work on the y-fastest twins for their x-principal faces; the per-warp
  // kernel does not, so a pass it handles alone needs neither the transposes nor the twin of the rate grid
  const bool run_cta = nwork > 0 || novf > 0;
  // The twins cost three grid-sized passes per call (transpose of tau_cell, clear and fold-back of the rate twin:
  // ~48 B per cell); without them the x-principal faces (a third of the updates) touch a 32-byte sector per 8-byte
  // value (~48 B per such update more).  So the twins pay from about 3 updates per mesh cell on -- always with
  // thousands of sources, never for the single-source and ten-source configurations.
  bool use_twins = est_updates >= 3.0 * (double)h->ncell;
  if (const char* env = getenv("C2B_TWINS")) use_twins = atoi(env) != 0;   // development / tests: force either way
  rp.use_twins = use_twins ? 1 : 0;
  const bool need_twins = use_twins && (run_cta || nwork_cl > 0);
  if (need_twins && h->taucell_t_dirty) {
    launch_to_yfast(h->d_taucell, h->d_taucell_t, c.mesh, h->stream);
    h->launches += 1;
    h->taucell_t_dirty = false;
  }
  if (need_twins) {
    CU(h, cudaMemsetAsync(h->d_phih_t, 0, h->ncell * sizeof(double), h->stream));
    if (h->d_phiheat_t) CU(h, cudaMemsetAsync(h->d_phiheat_t, 0, h->ncell * sizeof(double), h->stream));
  }
  CU(h, cudaMemsetAsync(h->d_ticket, 0, (novf > 0 ? 3 : 4) * sizeof(unsigned int), h->stream));
  CU(h, cudaEventRecord(h->ev[e0], h->stream));
  if (nwork_cl > 0) {  // long traces first: one cluster per source, planes in shared memory
    RtParams rc = rp;
    rc.work = d_work_cl;
    rc.nwork = nwork_cl;
    rc.ticket = h->d_ticket + 1;
    rc.scratch = h->d_scratch + raytrace_scratch_doubles_per_cta(h->plane_stride) * (size_t)h->rt.grid_cta;
    const int ncl = std::min(h->rt.clusters, nwork_cl);
    if (launch_raytrace_cluster(rc, ncl, h->stream)) {
      CU(h, cudaGetLastError());
      return fail(h, "cluster launch failed");
    }
    h->launches += 1;
  }
  if (nwork_w > 0 && novf == 0) {   // first subbox of the short traces, one warp per source
    RtParams rw = rp;
    rw.work = d_work_w;
    rw.nwork = nwork_w;
    rw.ticket = h->d_ticket + 2;
    rw.ovf_count = h->d_ticket + 3;
    launch_raytrace_warp(rw, std::min(h->rt.grid_warp, (nwork_w + h->rt.warp_warps - 1) / h->rt.warp_warps), h->rt.warp_warps, h->stream);
    h->launches += 1;
    CU(h, cudaGetLastError());
  }
  if (run_cta) {
    // the single-CTA kernel also takes what the per-warp kernel handed over (in this call or, novf > 0, in the previous one)
    if (nwork_w > 0 || novf > 0) rp.ovf_count = h->d_ticket + 3;
    const int grid = std::min(h->rt.grid_cta, nwork + (novf > 0 ? novf : nwork_w));
    launch_raytrace(rp, grid, h->stream);
    h->launches += 1;
    CU(h, cudaGetLastError());
  }
  if (need_twins) {
    launch_add_from_yfast(h->d_phih, h->d_phih_t, c.mesh, h->stream);
    h->launches += 1;
    if (h->d_phiheat) {
      launch_add_from_yfast(h->d_phiheat, h->d_phiheat_t, c.mesh, h->stream);
      h->launches += 1;
    }
    CU(h, cudaGetLastError());
  }
  CU(h, cudaEventRecord(h->ev[e0 + 1], h->stream));
  return 0;
}

static int64_t box_updates(const c2b_handle* h, int nbox) {
  if (nbox <= 0) return 0;
  int64_t u = 1;
  for (int d = 0; d < 3; ++d) {
    const int r = std::min(h->cfg.subboxsize * nbox, h->lim[d][1]);
    const int l = std::min(h->cfg.subboxsize * nbox, h->lim[d][0]);
    u *= (int64_t)(r + l + 1);
  }
  return u;
}

int c2b_begin_step(c2b_handle* h, double* sum_xh) {
  C2B_CHECK_H(h);
  if (int rc = check_ready(h)) return rc;
  if (bind_device(h)) return 1;
  ChemParams cp;
  fill_chem(h, 0.0, cp);
  // state_before(xh), photonstatistics.F90:104-132 (+ sum(xh) for evolve.F90:183)
  launch_stats(cp, h->d_xh, nullptr, h->chem_blocks, h->stream);
  h->launches += 1;
  if (int rc = fetch_stats(h)) return rc;
  h->h0_before = h->h_stats[kH0] * h->vol;
  h->h1_before = h->h_stats[kH1] * h->vol;
  h->sum_xh_intermed = h->h_stats[kSumXh];
  if (sum_xh) *sum_xh = h->sum_xh_intermed;
  // xh_av=xh ; xh_intermed=xh  (evolve.F90:140-147)
  CU(h, cudaMemcpyAsync(h->d_xh_av, h->d_xh, h->ncell * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_xh_intermed, h->d_xh, h->ncell * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  h->taucell_dirty = true;
  return 0;
}

int c2b_pass_all_sources(c2b_handle* h, int32_t niter, double dt, c2b_pass_report* rep) {
  C2B_CHECK_H(h);
  (void)dt;
  if (int rc = check_ready(h)) return rc;
  if (bind_device(h)) return 1;
  h->iter_niter = niter;   // what write_iteration_dump would record after this pass (evolve.F90:309)
  // set_rates_to_zero, evolve.F90:430-440
  CU(h, cudaMemsetAsync(h->d_phih, 0, h->ncell * sizeof(double), h->stream));
  if (h->d_phiheat) CU(h, cudaMemsetAsync(h->d_phiheat, 0, h->ncell * sizeof(double), h->stream));   // :435
  h->photon_loss = 0.0;
  h->LLS_loss = 0.0;
  float ms_rt = 0.f, ms_ar = 0.f;
  double loss_sum = 0.0, nbox_sum = 0.0, upd_sum = 0.0;
  if (h->NumSrc > 0) {
    CU(h, cudaMemsetAsync(h->d_nbox, 0, (size_t)h->NumSrc * sizeof(int), h->stream));
    CU(h, cudaMemsetAsync(h->d_loss, 0, (size_t)h->NumSrc * sizeof(double), h->stream));
    // route by the subbox count of the previous trace, longest first (the work queue is dynamic, so
    // the order only affects load balance and the order of the atomic additions)
    // A long trace parallelises over a cluster of 8 CTAs (one per octant); with many sources one CTA per
    // source already fills the GPU and has less synchronisation, so the cluster kernel is used only
    // while the long traces are few.  Sources never traced before (prediction 0) count as long when the
    // whole list is short.
    if (!h->routes_valid) {
    std::vector<int> small, large, tiny;
    const bool few = (int)h->work.size() <= h->cluster_max_sources;
    const bool zorder = !(getenv("C2B_NO_ZORDER") && atoi(getenv("C2B_NO_ZORDER")));
    for (int w : (zorder ? h->work_morton : h->work)) {
      const int pred = h->nbox_pred[w];
      const bool is_large = pred >= h->cluster_min_nbox || (pred == 0 && few && h->cluster_min_nbox < 100000);
      (is_large ? large : small).push_back(w);
    }
    if ((int)large.size() > h->cluster_max_sources) {
      small.insert(small.end(), large.begin(), large.end());
      large.clear();
    }
    // longest predicted trace first (counting sort on nbox, stable within a bucket)
    auto by_nbox_desc = [&](std::vector<int>& v) {
      int mx = 0;
      for (int w : v) mx = std::max(mx, h->nbox_pred[w]);
      std::vector<std::vector<int>> bucket((size_t)mx + 1);
      for (int w : v) bucket[(size_t)h->nbox_pred[w]].push_back(w);
      v.clear();
      for (int b = mx; b >= 0; --b) v.insert(v.end(), bucket[(size_t)b].begin(), bucket[(size_t)b].end());
    };
    // traces that ended after one subbox last time: one warp per source, if there are enough of them to fill the GPU
    if (h->rt.warp_warps > 0 && !(getenv("C2B_NO_WARP_KERNEL") && atoi(getenv("C2B_NO_WARP_KERNEL")))) {
      int n1 = 0;
      for (int w : small) n1 += h->nbox_pred[w] == 1;
      if (n1 >= h->warp_min_sources) {
        std::vector<int> rest;
        for (int w : small) (h->nbox_pred[w] == 1 ? tiny : rest).push_back(w);
        small.swap(rest);
      }
    }
    by_nbox_desc(large);
    by_nbox_desc(small);
    if (!tiny.empty())
      CU(h, cudaMemcpyAsync(h->d_work3, tiny.data(), tiny.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (!small.empty())
      CU(h, cudaMemcpyAsync(h->d_work, small.data(), small.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (!large.empty())
      CU(h, cudaMemcpyAsync(h->d_work2, large.data(), large.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    h->n_small = (int)small.size(); h->n_large = (int)large.size(); h->n_tiny = (int)tiny.size();
    // predicted updates of the CTA and cluster kernels (a source never traced before counts as a full box)
    h->est_updates = 0.0;
    for (const std::vector<int>* v : {&small, &large})
      for (int w : *v) {
        const int nb = h->nbox_pred[w];
        if ((size_t)nb >= h->updates_of_nbox.size())
          for (int k = (int)h->updates_of_nbox.size(); k <= nb; ++k) h->updates_of_nbox.push_back(box_updates(h, k));
        h->est_updates += nb > 0 ? (double)h->updates_of_nbox[(size_t)nb] : (double)h->ncell;
      }
    h->routes_valid = true;
    }
    const int n_small = h->n_small, n_large = h->n_large, n_tiny = h->n_tiny;
    if (int rc = trace_sources(h, h->d_work, n_small, h->d_work2, n_large, h->d_work3, n_tiny, 0, nullptr, 0, h->est_updates)) return rc;
    CU(h, cudaMemcpyAsync(h->h_nbox, h->d_nbox, (size_t)h->NumSrc * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(h->h_loss, h->d_loss, (size_t)h->NumSrc * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(h->h_ovf, h->d_ticket + 3, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaEventElapsedTime(&ms_rt, h->ev[0], h->ev[1]));
    if (n_small == 0 && n_tiny > 0 && *h->h_ovf > 0u) {
      // only the per-warp kernel ran and some of its sources need more than one subbox: the single-CTA kernel
      // carries on with those (when the single-CTA kernel runs in the same pass it takes them there and then)
      if (int rc = trace_sources(h, nullptr, 0, nullptr, 0, nullptr, 0, (int)*h->h_ovf, nullptr, 4,
                                 (double)*h->h_ovf * (double)box_updates(h, 2))) return rc;
      CU(h, cudaMemcpyAsync(h->h_nbox, h->d_nbox, (size_t)h->NumSrc * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CU(h, cudaMemcpyAsync(h->h_loss, h->d_loss, (size_t)h->NumSrc * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      CU(h, cudaStreamSynchronize(h->stream));
      float ms2 = 0.f;
      CU(h, cudaEventElapsedTime(&ms2, h->ev[4], h->ev[5]));
      ms_rt += ms2;
    }
    h->route_counts[0] += (long long)n_small;
    h->route_counts[1] += (long long)n_large;
    h->route_counts[2] += (long long)n_tiny;
    h->route_counts[3] += (long long)*h->h_ovf;
    // photon_loss(1)=photon_loss(1)+photon_loss_src ; sum_nbox=sum_nbox+nbox, in source order
    for (int w : h->work) {
      const int nb = h->h_nbox[w];
      if (!h->balance && nb != h->nbox_pred[w]) {   // (with several ranks: below, from the all-reduced counts)
        h->nbox_pred[w] = nb;
        h->routes_valid = false;   // the next pass sorts and routes again
      }
      loss_sum = loss_sum + h->h_loss[w];
      nbox_sum += (double)nb;
      if ((size_t)nb >= h->updates_of_nbox.size())
        for (int k = (int)h->updates_of_nbox.size(); k <= nb; ++k) h->updates_of_nbox.push_back(box_updates(h, k));
      upd_sum += (double)h->updates_of_nbox[(size_t)nb];
    }
  }
  if (h->cfg.nranks > 1) {
    // mpi_accumulate_grid_quantities, evolve.F90:577-616
    CU(h, cudaEventRecord(h->ev[2], h->stream));
    const int nr = h->cfg.nranks;
    const size_t nsmall = 4 + 2 * (size_t)nr;
    h->h_small[0] = loss_sum; h->h_small[1] = nbox_sum; h->h_small[2] = upd_sum; h->h_small[3] = 0.0;
    for (int r = 0; r < 2 * nr; ++r) h->h_small[4 + r] = 0.0;
    h->h_small[4 + 2 * h->cfg.rank] = (double)ms_rt;       // how long this rank traced ...
    h->h_small[5 + 2 * h->cfg.rank] = upd_sum;             // ... how many updates
    CU(h, cudaMemcpyAsync(h->d_small, h->h_small, nsmall * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    NC(h, g_nccl.AllReduce(h->d_phih, h->d_phih, h->ncell, ncclDouble, ncclSum, h->comm, h->stream));
    if (h->d_phiheat)   // evolve.F90:604-609
      NC(h, g_nccl.AllReduce(h->d_phiheat, h->d_phiheat, h->ncell, ncclDouble, ncclSum, h->comm, h->stream));
    NC(h, g_nccl.AllReduce(h->d_small, h->d_small, nsmall, ncclDouble, ncclSum, h->comm, h->stream));
    CU(h, cudaMemcpyAsync(h->h_small, h->d_small, nsmall * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (h->balance && h->NumSrc > 0) {
      // every rank learns the subbox count of every source (each source was traced by exactly one rank)
      NC(h, g_nccl.AllReduce(h->d_nbox, h->d_nbox, (size_t)h->NumSrc, ncclInt32, ncclSum, h->comm, h->stream));
      CU(h, cudaMemcpyAsync(h->h_nbox_all, h->d_nbox, (size_t)h->NumSrc * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(h, cudaEventRecord(h->ev[3], h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaEventElapsedTime(&ms_ar, h->ev[2], h->ev[3]));
    if (h->balance && h->NumSrc > 0) {
      // Deal the sources again for the next pass when the shares of the predicted work have drifted from the ranks'
      // measured speeds by more than 1.5 %: the same inputs on every rank (all-reduced), so the same decision and
      // the same assignment everywhere.
      for (int s2 = 0; s2 < h->NumSrc; ++s2)
        if (h->nbox_pred[(size_t)s2] != h->h_nbox_all[s2]) {
          h->nbox_pred[(size_t)s2] = h->h_nbox_all[s2];
          if (h->owner[(size_t)s2] == h->cfg.rank) h->routes_valid = false;   // this rank sorts and routes again
        }
      // relative speeds (mean 1), moved 30 % of the way from the last value to this pass's measurement
      bool have_speed = true;
      double ssum = 0.0;
      std::vector<double> speed((size_t)nr, 0.0);
      for (int r = 0; r < nr; ++r) {
        const double ms = h->h_small[4 + 2 * r], up = h->h_small[5 + 2 * r];
        if (ms > 0.0 && up > 0.0) speed[(size_t)r] = up / ms;
        else have_speed = false;
        ssum += speed[(size_t)r];
      }
      if (have_speed)
        for (int r = 0; r < nr; ++r) h->rank_speed[(size_t)r] = 0.7 * h->rank_speed[(size_t)r] + 0.3 * speed[(size_t)r] * nr / ssum;
      std::vector<int64_t> cost((size_t)h->NumSrc);
      std::vector<double> load((size_t)nr, 0.0);
      double total = 0.0, stot = 0.0;
      for (int s2 = 0; s2 < h->NumSrc; ++s2) {
        const int nb = h->nbox_pred[(size_t)s2];
        if ((size_t)nb >= h->updates_of_nbox.size())
          for (int k = (int)h->updates_of_nbox.size(); k <= nb; ++k) h->updates_of_nbox.push_back(box_updates(h, k));
        cost[(size_t)s2] = h->updates_of_nbox[(size_t)nb];
        const double c1 = (double)std::max<int64_t>(cost[(size_t)s2], 1);
        load[(size_t)h->owner[(size_t)s2]] += c1;
        total += c1;
      }
      for (int r = 0; r < nr; ++r) stot += h->rank_speed[(size_t)r];
      bool redeal = false;
      for (int r = 0; r < nr; ++r) {
        const double target = total * h->rank_speed[(size_t)r] / stot;
        if (std::fabs(load[(size_t)r] - target) > 0.015 * target) redeal = true;
      }
      if (redeal) {
        std::vector<int32_t> own((size_t)h->NumSrc);
        c2b_deal_sources(h->NumSrc, cost.data(), nr, h->rank_speed.data(), own.data());
        h->owner.assign(own.begin(), own.end());
        rebuild_work(h);
        h->redeals += 1;
        if (h->cfg.rank == 0 && getenv("C2B_DEBUG_BALANCE")) {
          fprintf(stderr, "c2b: sources dealt again (%lld): speeds", h->redeals);
          for (int r = 0; r < nr; ++r) fprintf(stderr, " %.4f", h->rank_speed[(size_t)r]);
          fprintf(stderr, "\n");
        }
      }
    }
    loss_sum = h->h_small[0]; nbox_sum = h->h_small[1]; upd_sum = h->h_small[2];
  }
  h->photon_loss = loss_sum;
  h->iter_photon_loss_all = loss_sum;
  if (rep) {
    rep->photon_loss_all = loss_sum;
    rep->sum_nbox_all = (int64_t)llround(nbox_sum);
    rep->updates = (int64_t)llround(upd_sum);
    rep->ms_raytrace = ms_rt;
    rep->ms_allreduce = ms_ar;
  }
  return 0;
}

int c2b_global_pass(c2b_handle* h, double dt, c2b_global_report* rep) {
  C2B_CHECK_H(h);
  if (int rc = check_ready(h)) return rc;
  if (bind_device(h)) return 1;
  // photon_loss(:)=photon_loss_all(:)/(real(mesh(1))*real(mesh(2))*real(mesh(3))), evolve.F90:525
  const float m = (float)h->cfg.mesh[0] * (float)h->cfg.mesh[1] * (float)h->cfg.mesh[2];
  h->photon_loss = h->iter_photon_loss_all / (double)m;
  ChemParams cp;
  fill_chem(h, dt, cp);
  CU(h, cudaEventRecord(h->ev[0], h->stream));
  launch_chemistry(cp, h->chem_blocks, h->stream);
  h->launches += 1;
  h->taucell_dirty = false;  // the chemistry kernel wrote tau_cell from the new xh_av
  h->taucell_t_dirty = true;
  CU(h, cudaEventRecord(h->ev[1], h->stream));
  if (int rc = fetch_stats(h)) return rc;
  float ms = 0.f;
  CU(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
  absorb_after(h, dt);  // calculate_photon_statistics(dt,xh_intermed,xh_av), evolve.F90:570
  h->sum_xh_intermed = h->h_stats[kSumXh];
  if (rep) {
    rep->conv_flag = (int32_t)llround(h->h_stats[kConv]);
    rep->min_avg_neutral = 1.0 - h->h_stats[kMaxXhAv];
    rep->sum_xh_intermed = h->sum_xh_intermed;
    fill_photon_stats(h, dt, &rep->stats);
    rep->ms_chemistry = ms;
  }
  return 0;
}

int c2b_end_step(c2b_handle* h, double dt, int32_t converged, c2b_photon_stats* final_stats) {
  C2B_CHECK_H(h);
  if (int rc = check_ready(h)) return rc;
  if (bind_device(h)) return 1;
  if (converged) {  // xh(:,:,:)=xh_intermed(:,:,:), evolve.F90:215-217
    CU(h, cudaMemcpyAsync(h->d_xh, h->d_xh_intermed, h->ncell * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (h->d_Tcur)  // set_final_temperature_point (:218, temperature_module.F90:173-183): current=intermed
      CU(h, cudaMemcpyAsync(h->d_Tcur, h->d_Tint, h->ncell * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  }
  // calculate_photon_statistics(dt,xh,xh_av), evolve.F90:277
  ChemParams cp;
  fill_chem(h, dt, cp);
  launch_stats(cp, h->d_xh, h->d_xh_av, h->chem_blocks, h->stream);
  h->launches += 1;
  if (int rc = fetch_stats(h)) return rc;
  absorb_after(h, dt);
  if (final_stats) fill_photon_stats(h, dt, final_stats);
  // update_grandtotal_photonstatistics, photonstatistics.F90:286-293
  h->grtotal_src = h->grtotal_src + h->sum_normflux * h->S_star * dt;
  h->grtotal_ion = h->grtotal_ion + h->total_ion - h->totcoll;
  return 0;
}

int c2b_evolve3d(c2b_handle* h, double time, double dt, int32_t restart, c2b_step_report* rep) {
  C2B_CHECK_H(h);
  (void)time;
  if (!rep) return fail(h, "c2b_evolve3d: null report");
  if (!(dt > 0)) return fail(h, "c2b_evolve3d: dt must be positive");
  memset(rep, 0, sizeof(*rep));
  const c2b_config& c = h->cfg;
  const long long launches0 = h->launches;
  if (bind_device(h)) return 1;
  CU(h, cudaEventRecord(h->ev_step[0], h->stream));
  int niter = 0;
  int conv_flag = c.mesh[0] * c.mesh[1] * c.mesh[2];
  // prev_sum_xh1_int=2.0*mesh(1)*mesh(2)*mesh(3): default-real arithmetic (evolve.F90:150-151)
  double prev_sum_xh1 = (double)(2.0f * (float)c.mesh[0] * (float)c.mesh[1] * (float)c.mesh[2]);
  double prev_sum_xh0 = prev_sum_xh1;
  double sum_xh1 = 0.0;
  int rc = 0;
  if (restart == 0) {
    rc = c2b_begin_step(h, &sum_xh1);
  } else {
    // start_from_dump + global_pass (evolve.F90:154-158); state_before still runs first (:136)
    if (!h->have_iter_state) rc = fail(h, "c2b_evolve3d: restart requested but c2b_set_iter_state was not called");
    if (!rc) {
      ChemParams cp;
      fill_chem(h, 0.0, cp);
      launch_stats(cp, h->d_xh, nullptr, h->chem_blocks, h->stream);
      h->launches += 1;
      rc = fetch_stats(h);
    }
    if (!rc) {
      h->h0_before = h->h_stats[kH0] * h->vol;
      h->h1_before = h->h_stats[kH1] * h->vol;
      // the restart branch does not touch prev_sum_xh*_int (evolve.F90:154-158): module variables without an
      // initialiser, zero in a freshly started run, which is when a restart happens
      prev_sum_xh1 = 0.0;
      prev_sum_xh0 = 0.0;
      niter = h->iter_niter;
      c2b_global_report gr;
      rc = c2b_global_pass(h, dt, &gr);
      conv_flag = gr.conv_flag;
      sum_xh1 = gr.sum_xh_intermed;
      rep->ms_chemistry += gr.ms_chemistry;
    }
  }
  if (rc) return rc;
  // conv_criterion=min(int(convergence_fraction*mesh(1)*mesh(2)*mesh(3)),(NumSrc-1)/3), evolve.F90:162
  const int c1 = (int)(c.convergence_fraction * (double)c.mesh[0] * (double)c.mesh[1] * (double)c.mesh[2]);
  const int c2 = (h->NumSrc - 1) / 3;
  const int conv_criterion = std::min(c1, c2);
  rep->conv_criterion = conv_criterion;
  int converged = 0;
  for (;;) {
    // sum_xh1_int=sum(xh_intermed) ; sum_xh0_int=real(N^3)-sum_xh1_int  (evolve.F90:183-196)
    const double sum_xh0 = (double)(float)(c.mesh[0] * c.mesh[1] * c.mesh[2]) - sum_xh1;
    const double rel1 = (sum_xh1 > 0.0) ? std::fabs(sum_xh1 - prev_sum_xh1) / sum_xh1 : 1.0;
    const double rel0 = (sum_xh0 > 0.0) ? std::fabs(sum_xh0 - prev_sum_xh0) / sum_xh0 : 1.0;
    if (niter < C2B_MAX_ITER) {
      rep->rel_change_sum_xh1[niter] = rel1;
      rep->rel_change_sum_xh0[niter] = rel0;
    }
    if (conv_flag < conv_criterion || (rel1 < c.convergence_fraction && rel0 < c.convergence_fraction)) {
      converged = 1;  // "Multiple sources convergence reached", :212-226
      break;
    } else if (niter > c.max_outer_iter) {
      converged = 0;  // 'Multiple sources not converging', :228-232
      break;
    }
    prev_sum_xh1 = sum_xh1;
    prev_sum_xh0 = sum_xh0;
    niter += 1;
    c2b_pass_report pr;
    if ((rc = c2b_pass_all_sources(h, niter, dt, &pr))) return rc;
    c2b_global_report gr;
    if ((rc = c2b_global_pass(h, dt, &gr))) return rc;
    conv_flag = gr.conv_flag;
    sum_xh1 = gr.sum_xh_intermed;
    h->iter_niter = niter;
    rep->ms_raytrace += pr.ms_raytrace;
    rep->ms_allreduce += pr.ms_allreduce;
    rep->ms_chemistry += gr.ms_chemistry;
    rep->total_updates += pr.updates;
    if (niter < C2B_MAX_ITER) {
      rep->conv_flag[niter] = conv_flag;
      rep->photon_loss_all[niter] = pr.photon_loss_all;
      rep->sum_nbox_all[niter] = pr.sum_nbox_all;
      rep->updates[niter] = pr.updates;
      rep->iter_stats[niter] = gr.stats;
    }
  }
  rep->niter = niter;
  rep->converged = converged;
  if ((rc = c2b_end_step(h, dt, converged, &rep->final_stats))) return rc;
  rep->grtotal_ion = h->grtotal_ion;
  rep->grtotal_src = h->grtotal_src;
  h->have_iter_state = false;
  CU(h, cudaEventRecord(h->ev_step[1], h->stream));
  CU(h, cudaEventSynchronize(h->ev_step[1]));
  float ms = 0.f;
  CU(h, cudaEventElapsedTime(&ms, h->ev_step[0], h->ev_step[1]));
  rep->ms_total = ms;
  rep->kernel_launches = h->launches - launches0;
  return 0;
}

// ---- outputs -------------------------------------------------------------------------------------
static int download(c2b_handle* h, void* dst, const void* src, size_t bytes, const char* what) {
  if (!dst) {
    h->error = std::string(what) + ": null pointer";
    return 3;
  }
  if (bind_device(h)) return 1;
  CU(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}
int c2b_get_xh(c2b_handle* h, double* p) { C2B_CHECK_H(h); return download(h, p, h->d_xh, h->ncell * 8, "c2b_get_xh"); }
int c2b_get_xh_av(c2b_handle* h, double* p) { C2B_CHECK_H(h); return download(h, p, h->d_xh_av, h->ncell * 8, "c2b_get_xh_av"); }
int c2b_get_xh_intermed(c2b_handle* h, double* p) { C2B_CHECK_H(h); return download(h, p, h->d_xh_intermed, h->ncell * 8, "c2b_get_xh_intermed"); }
int c2b_get_phih(c2b_handle* h, double* p) { C2B_CHECK_H(h); return download(h, p, h->d_phih, h->ncell * 8, "c2b_get_phih"); }
int c2b_get_phiheat(c2b_handle* h, double* p) {
  C2B_CHECK_H(h);
  if (int rc = need_thermal(h, "c2b_get_phiheat")) return rc;
  return download(h, p, h->d_phiheat, h->ncell * 8, "c2b_get_phiheat");
}
int c2b_get_iter_state_thermal(c2b_handle* h, double* phiheat, float* tg) {
  C2B_CHECK_H(h);
  if (int rc = need_thermal(h, "c2b_get_iter_state_thermal")) return rc;
  int rc = 0;
  if (phiheat && (rc = c2b_get_phiheat(h, phiheat))) return rc;
  if (tg && (rc = c2b_get_temperature_grid(h, tg))) return rc;
  return 0;
}
int c2b_set_iter_state_thermal(c2b_handle* h, const double* phiheat, const float* tg) {
  C2B_CHECK_H(h);
  if (int rc = need_thermal(h, "c2b_set_iter_state_thermal")) return rc;
  if (!phiheat || !tg) return fail(h, "c2b_set_iter_state_thermal: null pointer");
  if (bind_device(h)) return 1;
  CU(h, cudaMemcpyAsync(h->d_phiheat, phiheat, h->ncell * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return c2b_set_temperature_grid(h, tg);
}
int c2b_get_phih_f32(c2b_handle* h, float* p) {
  C2B_CHECK_H(h);
  if (!p) return fail(h, "c2b_get_phih_f32: null pointer");
  if (bind_device(h)) return 1;
  if (!h->d_f32tmp) CU(h, cudaMalloc(&h->d_f32tmp, h->ncell * sizeof(float)));
  launch_to_f32(h->d_phih, h->d_f32tmp, h->ncell, h->stream);
  h->launches += 1;
  return download(h, p, h->d_f32tmp, h->ncell * 4, "c2b_get_phih_f32");
}
int c2b_get_source_nbox(c2b_handle* h, int32_t* nbox) {
  C2B_CHECK_H(h);
  if (!nbox) return fail(h, "c2b_get_source_nbox: null pointer");
  // with several ranks and the load balance on, every rank holds the all-reduced counts of all sources
  if (h->NumSrc > 0) memcpy(nbox, h->balance ? h->h_nbox_all : h->h_nbox, (size_t)h->NumSrc * sizeof(int));
  return 0;
}
int c2b_get_source_loss(c2b_handle* h, double* loss) {
  C2B_CHECK_H(h);
  if (!loss) return fail(h, "c2b_get_source_loss: null pointer");
  if (h->NumSrc > 0) memcpy(loss, h->h_loss, (size_t)h->NumSrc * sizeof(double));
  return 0;
}

int c2b_get_iter_state(c2b_handle* h, int32_t* niter, double* photon_loss_all, double* phih, double* xh_av,
                       double* xh_intermed) {
  C2B_CHECK_H(h);
  if (niter) *niter = h->iter_niter;
  if (photon_loss_all) *photon_loss_all = h->iter_photon_loss_all;
  int rc = 0;
  if (phih && (rc = c2b_get_phih(h, phih))) return rc;
  if (xh_av && (rc = c2b_get_xh_av(h, xh_av))) return rc;
  if (xh_intermed && (rc = c2b_get_xh_intermed(h, xh_intermed))) return rc;
  return 0;
}

int c2b_set_iter_state(c2b_handle* h, int32_t niter, double photon_loss_all, const double* phih,
                       const double* xh_av, const double* xh_intermed) {
  C2B_CHECK_H(h);
  if (!phih || !xh_av || !xh_intermed) return fail(h, "c2b_set_iter_state: null pointer");
  if (bind_device(h)) return 1;
  const size_t b = h->ncell * sizeof(double);
  CU(h, cudaMemcpyAsync(h->d_phih, phih, b, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_xh_av, xh_av, b, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemcpyAsync(h->d_xh_intermed, xh_intermed, b, cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  h->iter_niter = niter;
  h->iter_photon_loss_all = photon_loss_all;
  h->have_iter_state = true;
  h->taucell_dirty = true;
  return 0;
}

void* c2b_dev_ptr(c2b_handle* h, const char* name) {
  if (!h || !name) return nullptr;
  if (!strcmp(name, "ndens")) return h->d_ndens;
  if (!strcmp(name, "xh")) return h->d_xh;
  if (!strcmp(name, "xh_av")) return h->d_xh_av;
  if (!strcmp(name, "xh_intermed")) return h->d_xh_intermed;
  if (!strcmp(name, "phih")) return h->d_phih;
  return nullptr;
}

// harness: keeps a device copy of xh so that a benchmark can start every step from the same state
int c2b_save_xh_dev(c2b_handle* h) {
  C2B_CHECK_H(h);
  if (!h->have_xh) return fail(h, "c2b_save_xh_dev: ionization fractions not set");
  if (bind_device(h)) return 1;
  if (!h->d_xh_saved) CU(h, cudaMalloc(&h->d_xh_saved, h->ncell * sizeof(double)));
  CU(h, cudaMemcpyAsync(h->d_xh_saved, h->d_xh, h->ncell * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int c2b_restore_xh_dev(c2b_handle* h) {
  C2B_CHECK_H(h);
  if (!h->d_xh_saved) return fail(h, "c2b_restore_xh_dev: nothing saved");
  if (bind_device(h)) return 1;
  CU(h, cudaMemcpyAsync(h->d_xh, h->d_xh_saved, h->ncell * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}

int c2b_synchronize(c2b_handle* h) {
  C2B_CHECK_H(h);
  if (bind_device(h)) return 1;
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int c2b_trace_source_debug(c2b_handle* h, int32_t ns, double* coldensh_out, double* phih, int32_t* nbox,
                           double* photon_loss_src) {
  C2B_CHECK_H(h);
  if (int rc = check_ready(h)) return rc;
  if (ns < 1 || ns > h->NumSrc) return fail(h, "c2b_trace_source_debug: source number out of range");
  if (bind_device(h)) return 1;
  double* d_dbg = nullptr;
  int* d_one = nullptr;
  CU(h, cudaMalloc(&d_dbg, h->ncell * sizeof(double)));
  CU(h, cudaMalloc(&d_one, sizeof(int)));
  const int idx = ns - 1;
  CU(h, cudaMemcpyAsync(d_one, &idx, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaMemsetAsync(d_dbg, 0, h->ncell * sizeof(double), h->stream));
  CU(h, cudaMemsetAsync(h->d_phih, 0, h->ncell * sizeof(double), h->stream));
  // the trace reads xh_av; outside evolve3D that is the caller's responsibility (tests copy xh)
  // C2B_DEBUG_CLUSTER=1 sends the diagnostic trace through the cluster kernel
  const char* envc = getenv("C2B_DEBUG_CLUSTER");
  const bool use_cl = envc && atoi(envc) != 0;
  int rc = use_cl ? trace_sources(h, nullptr, 0, d_one, 1, nullptr, 0, 0, d_dbg, 0) : trace_sources(h, d_one, 1, nullptr, 0, nullptr, 0, 0, d_dbg, 0);
  if (!rc && coldensh_out) rc = download(h, coldensh_out, d_dbg, h->ncell * 8, "coldensh_out");
  if (!rc && phih) rc = download(h, phih, h->d_phih, h->ncell * 8, "phih");
  int nb = 0;
  double ls = 0.0;
  if (!rc) rc = download(h, &nb, h->d_nbox + idx, sizeof(int), "nbox");
  if (!rc) rc = download(h, &ls, h->d_loss + idx, sizeof(double), "loss");
  if (nbox) *nbox = nb;
  if (photon_loss_src) *photon_loss_src = ls;
  cudaFree(d_dbg);
  cudaFree(d_one);
  return rc;
}

int c2b_get_route_counts(c2b_handle* h, int64_t counts[4]) {
  C2B_CHECK_H(h);
  if (!counts) return fail(h, "c2b_get_route_counts: null argument");
  for (int i = 0; i < 4; ++i) counts[i] = (int64_t)h->route_counts[i];
  return 0;
}

int c2b_get_source_owner(c2b_handle* h, int32_t* owner) {
  C2B_CHECK_H(h);
  if (!owner && h->NumSrc > 0) return fail(h, "c2b_get_source_owner: null argument");
  for (int s = 0; s < h->NumSrc; ++s) owner[s] = h->owner[(size_t)s];
  return 0;
}

int c2b_measure_dfma_rate(c2b_handle* h, double* dfma_per_s) {
  C2B_CHECK_H(h);
  if (!dfma_per_s) return fail(h, "c2b_measure_dfma_rate: null pointer");
  if (bind_device(h)) return 1;
  *dfma_per_s = measure_dfma_rate(h->stream);
  h->launches += 4;
  CU(h, cudaGetLastError());
  return 0;
}

}  // extern "C"
