// Short-characteristics ray tracer of C2-Ray3Dm as a Chebyshev-shell wavefront kernel (sm_100a).
//
// Replaces do_source / evolve2D / evolve1D_axis / evolve2D_plane / evolve3D_quadrant
// (evolve_source.F90:58-591), evolve0D (evolve_point.F90:83-299), cinterp + weightf
// (column_density.f90:29-293), coldens (doric.f90:141-155) and photoion_rates with its table
// look-up (radiation_photoionrates.F90:71-317).
//
// Decomposition.  The reference walks the cells of a growing cubic subbox in an order that
// guarantees the four upstream neighbours of a cell are finished first.  Those neighbours are
// always one step closer to the source along every axis, and the ones that are not in the previous
// Chebyshev shell carry an interpolation weight of exactly 0, so shell r depends on
// shell r-1 only.  A shell further splits into 24 independent "face quadrants": principal axis p
// (the dominant |offset|, which selects the cinterp branch), the sign of the principal offset and
// the signs of the two transverse offsets.  In quadrant-local coordinates (a,b) = transverse
// distances, the cell (a,b) of plane r reads (a-1|a, b-1|b) of plane r-1 of the SAME quadrant.
// Cells shared between quadrants (on-axis a==0 / b==0, cube edges a==r / b==r) are computed by
// every quadrant that needs them as an upstream value (the three cinterp branches give the same
// value on ties), but only the quadrant the reference's branch order selects "owns" the cell:
// it alone adds the rate into phih_grid, counts the boundary loss and the update.
//
// v0 work distribution: one CTA per source, persistent CTAs pulling sources from an atomic ticket
// (the device-side do_grid_master, master_slave.F90:124-231); the 24 planes of the current and
// previous shell live in a per-CTA global scratch that stays L2 resident for moderate radii.
//
// This translation unit is compiled with -fmad=false: the interpolation weights must evaluate to
// exact zeros where the reference's do (SURVEY hard part 4), and it keeps the arithmetic within
// rounding of the CPU restatement.
#include "c2b_common.cuh"

namespace c2b {
namespace {

constexpr int kThreads = 256;
constexpr int kQuadrants = 24;

__device__ __forceinline__ double weightf(double cd, double sig) {
  // column_density.f90:276-293
  return 1.0 / fmax(0.6, cd * sig);
}

struct TablePos {
  double residual;
  int ipos, ipos_p1;
};

// set_tau_table_positions, radiation_photoionrates.F90:184-208
__device__ __forceinline__ TablePos table_pos(double tau, double minlogtau, double dlogtau) {
  TablePos t;
  double lt = log10(fmax(1.0e-20, tau));
  double od = fmin((double)kNumTau, fmax(0.0, 1.0 + (lt - minlogtau) / dlogtau));
  t.ipos = (int)od;
  t.residual = od - (double)t.ipos;
  t.ipos_p1 = min(kNumTau, t.ipos + 1);
  return t;
}

// read_table, radiation_photoionrates.F90:212-228
__device__ __forceinline__ double read_table(const double* tab, const TablePos& t) {
  double lo = tab[t.ipos];
  return lo + (tab[t.ipos_p1] - lo) * t.residual;
}

__device__ __forceinline__ int wrap(int x, int n) {
  // modulo(x-1,mesh)+1 of evolve_point.F90:122-124 for |offset| <= n/2, 0-based
  if (x < 0) x += n;
  else if (x >= n) x -= n;
  return x;
}

__global__ void __launch_bounds__(kThreads, 2) raytrace_kernel(RtParams P) {
  __shared__ double s_thick[kTableLen];
  __shared__ double s_thin[kTableLen];
  __shared__ double s_red[kThreads / 32];
  __shared__ double s_loss;
  __shared__ int s_work;

  const int tid = threadIdx.x;
  for (int i = tid; i < kTableLen; i += kThreads) {
    s_thick[i] = P.thick[i];
    s_thin[i] = P.thin[i];
  }
  const int S = P.plane_stride;
  const size_t plane_sz = (size_t)kQuadrants * S * S;
  double* buf0 = P.scratch + (size_t)blockIdx.x * 2 * plane_sz;
  double* buf1 = buf0 + plane_sz;
  const double dr0 = P.dr[0], dr1 = P.dr[1], dr2 = P.dr[2];

  for (;;) {
    __syncthreads();  // also orders the table fill and the previous source's last reads of s_work
    if (tid == 0) s_work = (int)atomicAdd(P.ticket, 1u);
    __syncthreads();
    const int w = s_work;
    if (w >= P.nwork) break;
    const int ns = P.work[w];  // 0-based source index
    // 1-based source position as the reference holds it; 0-based = minus one
    const int src1[3] = {P.srcpos[3 * ns], P.srcpos[3 * ns + 1], P.srcpos[3 * ns + 2]};
    const double normflux = P.normflux[ns];
    const double total_source_flux = normflux * P.S_star;  // evolve_source.F90:119

    double* prev = buf0;
    double* cur = buf1;
    int nbox = 0;
    double photon_loss_src = total_source_flux;  // :121
    int lr[3] = {0, 0, 0}, ll[3] = {0, 0, 0};    // last_r-src, src-last_l
    int r_done = -1;                             // last shell finished
    // do while (evolve_source.F90:128-131); all threads evaluate it on identical values
    while (photon_loss_src > P.loss_fraction * total_source_flux && lr[2] < P.lim[2][1] &&
           ll[2] < P.lim[2][0]) {
      nbox += 1;
      int rmax = 0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {  // :135-136
        lr[d] = min(P.subboxsize * nbox, P.lim[d][1]);
        ll[d] = min(P.subboxsize * nbox, P.lim[d][0]);
        rmax = max(rmax, max(lr[d], ll[d]));
      }
      double loss = 0.0;
      for (int r = r_done + 1; r <= rmax; ++r) {
        const int P1 = r + 1;
        const int per_q = P1 * P1;
        const int total = kQuadrants * per_q;
        const double rp = (double)r;
        const double alam = (rp - 0.5) / rp;  // (real(km-k0)+sgnk*0.5)/dk, column_density.f90:109
        for (int idx = tid; idx < total; idx += kThreads) {
          const int q = idx / per_q;
          const int rem = idx - q * per_q;
          const int b = rem / P1;
          const int a = rem - b * P1;
          const int p = q >> 3;  // 0: z principal, 1: y, 2: x  (branch order of cinterp)
          const int sp = (q & 4) ? -1 : 1, sa = (q & 2) ? -1 : 1, sb = (q & 1) ? -1 : 1;
          const int axP = (p == 0) ? 2 : (p == 1 ? 1 : 0);
          const int axA = (p == 2) ? 1 : 0;
          const int axB = (p == 0) ? 1 : 2;
          // box of this pass (static limits bind before 5*nbox does not matter: r <= 5*nbox)
          if (r > (sp > 0 ? lr[axP] : ll[axP]) || a > (sa > 0 ? lr[axA] : ll[axA]) ||
              b > (sb > 0 ? lr[axB] : ll[axB]))
            continue;
          int d[3];
          d[axP] = sp * r;
          d[axA] = sa * a;
          d[axB] = sb * b;
          const int i0 = wrap(src1[0] - 1 + d[0], P.n[0]);
          const int j0 = wrap(src1[1] - 1 + d[1], P.n[1]);
          const int k0 = wrap(src1[2] - 1 + d[2], P.n[2]);
          const size_t cell = ((size_t)k0 * P.n[1] + j0) * P.n[0] + i0;
          // ownership: the quadrant the reference's branch order and sign(1,0)=+1 select
          bool owner = (a > 0 || sa > 0) && (b > 0 || sb > 0) && (r > 0 || sp > 0);
          if (p == 1) owner = owner && (b < r);
          if (p == 2) owner = owner && (a < r) && (b < r);
          if (r == 0) owner = owner && (p == 0);

          const double h_av1 = fmax(P.xh_av[cell], P.epsilon);   // evolve_point.F90:137
          const double h_av0 = fmax(1.0 - h_av1, P.epsilon);     // :140
          const double ndens_p = (double)P.ndens[cell];          // :145
          double coldensh_in, path, vol_ph;
          bool stop = false;
          if (r == 0) {  // :151-160
            coldensh_in = 0.0;
            path = 0.5 * dr0;
            vol_ph = dr0 * dr1 * dr2;
          } else {
            // cinterp in quadrant-local form (column_density.f90:108-171 and its y/x twins)
            const double dA = (double)d[axA], dB = (double)d[axB];
            const int sgA = (a == 0) ? 1 : sa, sgB = (b == 0) ? 1 : sb;  // sign(1,idel)
            const double xc = alam * dA + (double)src1[axA];
            const double yc = alam * dB + (double)src1[axB];
            const double amh = (double)(src1[axA] + d[axA] - sgA) + 0.5 * (double)sgA;
            const double bmh = (double)(src1[axB] + d[axB] - sgB) + 0.5 * (double)sgB;
            const double dx = 2.0 * fabs(xc - amh);
            const double dy = 2.0 * fabs(yc - bmh);
            const double s1 = (1.0 - dx) * (1.0 - dy);
            const double s2 = (1.0 - dy) * dx;
            const double s3 = (1.0 - dx) * dy;
            const double s4 = dx * dy;
            // upstream cells of plane r-1; the ones outside it have weight exactly 0
            const double* pl = prev + (size_t)q * S * S;
            const bool am_ok = a >= 1, a_ok = a <= r - 1, bm_ok = b >= 1, b_ok = b <= r - 1;
            const double c1 = (am_ok && bm_ok) ? pl[(b - 1) * S + (a - 1)] : 0.0;
            const double c2 = (a_ok && bm_ok) ? pl[(b - 1) * S + a] : 0.0;
            const double c3 = (am_ok && b_ok) ? pl[b * S + (a - 1)] : 0.0;
            const double c4 = (a_ok && b_ok) ? pl[b * S + a] : 0.0;
            const double w1 = s1 * weightf(c1, P.sigma_HI);
            const double w2 = s2 * weightf(c2, P.sigma_HI);
            const double w3 = s3 * weightf(c3, P.sigma_HI);
            const double w4 = s4 * weightf(c4, P.sigma_HI);
            double cdensi = (c1 * w1 + c2 * w2 + c3 * w3 + c4 * w4) / (w1 + w2 + w3 + w4);
            if (r == 1 && (a == 1 || b == 1)) {  // :152-158
              cdensi = ((a == 1 && b == 1) ? P.sqrt3 : P.sqrt2) * cdensi;
            }
            const double pathc = sqrt((dA * dA + dB * dB) / (rp * rp) + 1.0);
            coldensh_in = cdensi;
            path = pathc * dr0;                                   // evolve_point.F90:166
            const double xs = dr0 * (double)d[0];
            const double ys = dr1 * (double)d[1];
            const double zs = dr2 * (double)d[2];
            const double dist2 = xs * xs + ys * ys + zs * zs;
            vol_ph = 4.0 * P.pi * dist2 * path;                   // :177
            if (P.use_lls) {                                      // :186-196
              if (P.type_lls == 3) {
                if (dist2 > P.rmax_lls2) stop = true;
              } else {
                const double cl = (P.type_lls == 2) ? (double)P.lls_grid[cell] : P.coldensh_lls;
                coldensh_in = coldensh_in + cl * path / dr0;
              }
            }
          }
          if (coldensh_in > P.max_coldensh) stop = true;          // :201
          const double cd_out = coldensh_in + h_av0 * ndens_p * path;  // :247-248
          cur[(size_t)q * S * S + b * S + a] = cd_out;
          if (!owner) continue;
          if (P.coldens_dbg) P.coldens_dbg[cell] = cd_out;
          if (!stop && normflux > 0.0) {
            // photoion_rates / photo_lookuptable, radiation_photoionrates.F90:71-317
            const double tau_in = coldensh_in * P.sigma_HI;
            const double tau_out = cd_out * P.sigma_HI;
            const TablePos pin = table_pos(tau_in, P.minlogtau, P.dlogtau);
            const double phi_in = normflux * read_table(s_thick, pin);
            double phi_out, phi_all;
            if (fabs(tau_out - tau_in) > P.tau_photo_limit) {
              const TablePos pout = table_pos(tau_out, P.minlogtau, P.dlogtau);
              phi_out = normflux * read_table(s_thick, pout);
              phi_all = phi_in - phi_out;
            } else {
              phi_all = normflux * (tau_out - tau_in) * read_table(s_thin, pin);
              phi_out = phi_in - phi_all;
            }
            double photo_cell = phi_all / vol_ph;
            photo_cell = photo_cell / (h_av0 * ndens_p);          // evolve_point.F90:262
            if (photo_cell != 0.0) atomicAdd(&P.phih[cell], photo_cell);  // :283-284
            // boundary of this pass's subbox (:290-295)
            if (d[0] == -ll[0] || d[1] == -ll[1] || d[2] == -ll[2] || d[0] == lr[0] ||
                d[1] == lr[1] || d[2] == lr[2])
              loss = loss + phi_out * P.vol / vol_ph;
          }
        }
        __syncthreads();  // plane r complete before plane r+1 reads it
        double* t = prev;
        prev = cur;
        cur = t;
      }
      r_done = rmax;
      // photon_loss_src = sum over the CTA (plays photon_loss_src_thread, evolve_source.F90:183-186)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
      if ((tid & 31) == 0) s_red[tid >> 5] = loss;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < kThreads / 32; ++i) t += s_red[i];
        s_loss = t;
      }
      __syncthreads();
      photon_loss_src = s_loss;
    }
    if (tid == 0) {
      P.nbox_out[ns] = nbox;              // sum_nbox=sum_nbox+nbox, :219
      P.loss_out[ns] = photon_loss_src;   // photon_loss(1)=photon_loss(1)+photon_loss_src, :216
    }
  }
}

}  // namespace

size_t raytrace_scratch_doubles_per_cta(int plane_stride) {
  return (size_t)2 * kQuadrants * plane_stride * plane_stride;
}

int raytrace_max_grid() {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, raytrace_kernel, kThreads, 0);
  if (per_sm < 1) per_sm = 1;
  return sms * per_sm;
}

void launch_raytrace(const RtParams& p, int grid, cudaStream_t stream) {
  raytrace_kernel<<<grid, kThreads, 0, stream>>>(p);
}

}  // namespace c2b
