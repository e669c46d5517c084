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
// every quadrant that needs them as an upstream value (the three cinterp branches agree on ties),
// but only the quadrant the reference's branch order selects "owns" the cell: it alone adds the
// rate into phih_grid and counts the boundary loss.
//
// Mapping.  One work group per source, persistent work groups pulling sources from an atomic ticket
// (the device-side do_grid_master, master_slave.F90:124-231).  A work group is one CTA of 256 threads
// (three resident per SM) or, when there are too few long traces to fill the GPU that way, a
// thread-block cluster of 8 CTAs, one per octant with its three face quadrants, whose boundary-loss
// partial sums meet in rank 0's shared memory over DSMEM once per subbox pass.  Within a shell a
// thread owns a column (quadrant q, transverse index a), or a b-segment of it, and walks b: the two
// upstream values of its own column stay in registers from one b to the next, the two of column a-1
// arrive by warp shuffle, so a cell costs one plane load, requested one row ahead together with the
// grid value.  The planes of shell r-1 and r live in shared memory while they fit, afterwards in a
// per-CTA global scratch.  The optical-depth table is staged in shared memory as (value, forward
// difference) pairs.  Quadrants whose principal axis is x walk planes of constant x; they read and
// accumulate into y-fastest twins of the grids so that their accesses are contiguous too.
//
// Arithmetic.  The planes hold optical depths tau = sigma_HI * N_HI; the per-cell opacity
// tau_cell = sigma_HI*dr(1)*max(1-max(xh_av,eps),eps)*ndens comes from a grid the per-cell kernel
// writes, so an update reads 8 bytes and adds 8.  The reference's five divisions per interpolation
// collapse into one (common denominator), the two log10 of the table look-up become a 128-entry
// table + degree-5 polynomial log2 folded into the table coordinate, the path length comes from a
// reciprocal square root, and the reciprocals are hardware seeds with one third-order refinement.
// All of it stays within ~1e-13 of the CPU restatement (tests bound the rates at 1e-6 relative as
// BASELINE.json requires).
#include <cooperative_groups.h>

#include <algorithm>
#include <vector>

#include "c2b_common.cuh"

namespace cg = cooperative_groups;

namespace c2b {
namespace {

constexpr int kThreadsCta = 256;   // one CTA per source
constexpr int kQuadrants = 24;
#ifndef C2B_CTA_PER_SM
#define C2B_CTA_PER_SM 3
#endif
constexpr int kCtaPerSm = C2B_CTA_PER_SM;   // resident CTAs per SM of the one-CTA-per-source kernel

// max/min of two NON-NEGATIVE doubles through their bit patterns (integer order == numeric order there);
// avoids the NaN-propagating DSETP.MAX/FSEL/LOP3 sequence fmax() and ?: compile to.
__device__ __forceinline__ double pos_min(double x, double y) {
  return __longlong_as_double(min(__double_as_longlong(x), __double_as_longlong(y)));
}

// 1/x: hardware seed (>= 20 bits) refined with y*(1+e+e^2), e = 1-x*y  =>  relative error ~e^3 < 1e-17
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);
  return fma(t, y, y);
}

// 1/sqrt(x): hardware seed refined with y*(1 + e/2 + 3e^2/8), e = 1-x*y*y  =>  relative error ~e^3
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);
  const double t = fma(0.375, e, 0.5) * e;
  return fma(y, t, y);
}

// x > y ? x : y on the FP64 compare without fmax()'s NaN handling (3 instructions)
__device__ __forceinline__ double sel_max(double x, double y) {
  double r;
  asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(x), "d"(y));
  return r;
}

// table coordinate odpos = 1 + (log10(max(1e-20,tau)) - minlogtau)/dlogtau of
// set_tau_table_positions (radiation_photoionrates.F90:184-208), clamped to NumTau.
// logtab[j] = {1/c_j, A + B*log2(c_j)}, c_j = 1 + (j+0.5)/128; coef = B/ln2 * {1,-1/2,1/3,-1/4,1/5}
__device__ __forceinline__ double table_coord(double tau, const double2* __restrict__ logtab,
                                              const RtParams& P) {
  const double t = sel_max(tau, 1.0e-20);
  const int hi = __double2hiint(t);
  const int lo = __double2loint(t);
  const int e = (hi >> 20) - 1023;
  const int j = (hi >> 13) & 127;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const double2 lt = logtab[j];
  const double rr = fma(m, lt.x, -1.0);
  double q = fma(rr, P.logc[4], P.logc[3]);
  q = fma(rr, q, P.logc[2]);
  q = fma(rr, q, P.logc[1]);
  q = fma(rr, q, P.logc[0]);
  double od = fma(P.logB, (double)e, lt.y);
  od = fma(rr, q, od);
  return pos_min(od, (double)kNumTau);  // od >= 0 because tau >= 1e-20
}

// read_table (radiation_photoionrates.F90:212-228) on the (value, forward difference) pairs
__device__ __forceinline__ double lerp_pairs(const double2* __restrict__ tab, double od) {
  const int ipos = (int)od;
  const double res = od - (double)ipos;
  const double2 t = tab[ipos];
  return fma(t.y, res, t.x);
}

// photoion_rates / photo_lookuptable (radiation_photoionrates.F90:71-317) for one stellar source:
// Gamma_cell*vol_ph = F*(thick(tau_in)-thick(tau_out)), or F*dtau*thin(tau_in) below tau_photo_limit.
__device__ __forceinline__ void photo_rates(double tau_in, double tau_out, double normflux,
                                            const double2* __restrict__ s_thick, const double2* __restrict__ s_logtab,
                                            const RtParams& P, double& phi_all, double& phi_out) {
  const double od_in = table_coord(tau_in, s_logtab, P);
  const double phi_in = normflux * lerp_pairs(s_thick, od_in);
  const double dtau = tau_out - tau_in;
  if (fabs(dtau) > P.tau_photo_limit) {
    const double od_out = table_coord(tau_out, s_logtab, P);
    phi_out = normflux * lerp_pairs(s_thick, od_out);
    phi_all = phi_in - phi_out;
  } else {
    const int ipos = (int)od_in;
    const double res = od_in - (double)ipos;
    const double lo = P.thin[ipos];
    const double thin = lo + (P.thin[min(kNumTau, ipos + 1)] - lo) * res;
    phi_all = normflux * dtau * thin;
    phi_out = phi_in - phi_all;
  }
}

// dist2 = xs*xs+ys*ys+zs*zs of evolve_point.F90:170-174 for the cell (principal offset r, transverse a, b).
// Cubic cells (the reference's grids: dr(1)=dr(2)=dr(3)) need only dr^2 * (r^2+a^2+b^2) = dr^2 * q2; the general
// form is evaluated on demand so that no per-axis factors stay live in the row loop.
__device__ __forceinline__ double dist2_of(const RtParams& P, int p, int r, int a, double b2, double q2) {
  if (P.cubic_cells) return P.dr2[0] * q2;
  const double dP = (p == 0) ? P.dr2[2] : (p == 1 ? P.dr2[1] : P.dr2[0]);
  const double dA = (p == 2) ? P.dr2[1] : P.dr2[0];
  const double dB = (p == 0) ? P.dr2[1] : P.dr2[2];
  return fma(dB, b2, fma(dP, (double)(r * r), dA * (double)(a * a)));
}

__device__ __forceinline__ int wrap(int x, int n) {
  // modulo(x-1,mesh)+1 of evolve_point.F90:122-124 for |offset| <= n/2, 0-based
  if (x < 0) x += n;
  else if (x >= n) x -= n;
  return x;
}

// One CTA (kCluster == 1) or one cluster of 8 CTAs (kCluster == 8, one CTA per octant with its three
// face quadrants) per source.
// kLls: 0 = no LLS, 1 = homogeneous, 2 = LLS_grid, 3 = R_max barrier (LLS.F90:107-116); kDebug adds
// the coldensh_out diagnostic store.
// kGroups: the CTA's warps form kGroups independent groups, each owning kNq/kGroups quadrants and its own
// named barrier, so a group waiting for its shell to complete does not idle the others.
template <int kT, int kCluster, int kGroups, int kLls, bool kDebug>
__global__ void __launch_bounds__(kT, (kT <= 256) ? ((kCluster == 1) ? kCtaPerSm : 2) : 1) raytrace_kernel(RtParams P) {
  constexpr int kNq = kQuadrants / kCluster;        // face quadrants handled by this CTA
  constexpr int kNqg = kNq / kGroups;               // ... by one warp group
  constexpr int kTg = kT / kGroups;                 // threads per group
  static_assert(kNq % kGroups == 0 && kT % kGroups == 0 && kTg % 32 == 0, "bad group split");
  extern __shared__ double2 smem2[];
  double2* s_thick = smem2;                         // kTableLen pairs
  double2* s_logtab = smem2 + kTableLen;            // 128 pairs
  double* s_planes = reinterpret_cast<double*>(smem2 + kTableLen + 128);
  __shared__ double s_red[kT / 32];
  __shared__ double s_slot[2][8];                   // per-octant boundary loss, alternating by pass (rank 0's copy is used)
  __shared__ int s_work;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int grp = tid / kTg;
  const int gtid = tid - grp * kTg;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = (kCluster > 1) ? cluster.block_rank() : 0u;
  const unsigned cid = (kCluster > 1) ? (blockIdx.x / kCluster) : blockIdx.x;   // work-group index
  (void)cid;
  for (int i = tid; i < kTableLen; i += kT) s_thick[i] = P.thick2[i];
  for (int i = tid; i < 128; i += kT) s_logtab[i] = P.logtab[i];
  // per group: two shared plane buffers of `cap` doubles and two global ones of kNqg*S*S
  const int cap = (((kCluster == 1) ? P.smem_plane_doubles : P.smem_plane_doubles_cl) / kGroups) & ~1;
  double* s_planes_g = s_planes + (size_t)grp * 2 * cap;
  const size_t gplane = (size_t)kNqg * P.plane_stride * P.plane_stride;
  double* gbuf0 = P.scratch + ((size_t)blockIdx.x * kGroups + grp) * 2 * gplane;
  double* gbuf1 = gbuf0 + gplane;
  const int n0 = P.n[0], n1 = P.n[1], n2 = P.n[2];
  const unsigned st1 = (unsigned)n0, st2 = (unsigned)n0 * (unsigned)n1;   // x-fastest strides of y and z
  int pass_parity = 0;

  for (;;) {
    // ---- next source (device-side do_grid_master: one ticket per work group) ---------------------
    __syncthreads();
    if (crank == 0 && tid == 0) s_work = (int)atomicAdd(P.ticket, 1u);
    int w;
    if (kCluster > 1) {
      cluster.sync();
      w = *cluster.map_shared_rank(&s_work, 0);
    } else {
      __syncthreads();
      w = s_work;
    }
    if (w >= P.nwork) break;
    const int ns = P.work[w];  // 0-based source index
    const int src0 = P.srcpos[3 * ns] - 1, src1 = P.srcpos[3 * ns + 1] - 1, src2 = P.srcpos[3 * ns + 2] - 1;
    const double normflux = P.normflux[ns];
    const double total_source_flux = normflux * P.S_star;  // evolve_source.F90:119

    int nbox = 0;
    double photon_loss_src = total_source_flux;  // :121
    int lr0 = 0, lr1 = 0, lr2 = 0, ll0 = 0, ll1 = 0, ll2 = 0;  // last_r-src, src-last_l per axis
    int r_done = -1;
    // do while (evolve_source.F90:128-131); every thread of the work group evaluates it on identical values
    while (photon_loss_src > P.loss_fraction * total_source_flux && lr2 < P.lim[2][1] &&
           ll2 < P.lim[2][0]) {
      nbox += 1;
      const int reach = P.subboxsize * nbox;  // :135-136
      lr0 = min(reach, P.lim[0][1]); ll0 = min(reach, P.lim[0][0]);
      lr1 = min(reach, P.lim[1][1]); ll1 = min(reach, P.lim[1][0]);
      lr2 = min(reach, P.lim[2][1]); ll2 = min(reach, P.lim[2][0]);
      const int rmax = max(max(max(lr0, ll0), max(lr1, ll1)), max(lr2, ll2));
      double loss = 0.0;
      if (r_done < 0) {
        // shell 0 = the source cell (evolve_point.F90:151-160): coldensh_in=0, path=dr/2, vol_ph=cell volume.
        // Every quadrant of the group stores it as its plane 0; the (+,+,+) z quadrant owns it.
        if (gtid < kNqg) {
          const int qc = grp * kNqg + gtid;
          const int oct = (int)crank * (8 / kCluster) + qc / 3;
          const unsigned cell = (unsigned)src2 * st2 + (unsigned)src1 * st1 + (unsigned)src0;
          const double tau_cell = P.tau_cell[cell];
          const double tau_out = 0.5 * tau_cell;
          s_planes_g[gtid] = tau_out;   // plane 0 always fits in shared memory
          if (oct == 0 && qc - (qc / 3) * 3 == 0) {
            if (kDebug) P.coldens_dbg[cell] = tau_out * P.inv_sigma;
            if (normflux > 0.0) {
              double phi_all, phi_out;
              photo_rates(0.0, tau_out, normflux, s_thick, s_logtab, P, phi_all, phi_out);
              // rate = phi_all/(vol_cell*nHI), nHI = tau_cell/(sigma*dr0)
              const double photo_cell = phi_all * fast_rcp(P.vol_cell * tau_cell * P.inv_sigma_dr0);
              if (photo_cell != 0.0) atomicAdd(&P.phih[cell], photo_cell);
            }
          }
        }
        if (kGroups == 1) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(kTg) : "memory");
        r_done = 0;
      }
      for (int r = r_done + 1; r <= rmax; ++r) {
        const int P1 = r + 1;
        // plane buffers of shell r (cur) and r-1 (prev): shared while they fit, else global scratch
        double* cur = (kNqg * P1 * P1 <= cap) ? (s_planes_g + (r & 1) * cap) : ((r & 1) ? gbuf1 : gbuf0);
        const double* prev = (kNqg * r * r <= cap) ? (s_planes_g + ((r - 1) & 1) * cap)
                                                   : (((r - 1) & 1) ? gbuf1 : gbuf0);
        const double inv_r = 1.0 / (double)r;
        // work items: (segment of b, quadrant, column a), a fastest so that a warp spans adjacent columns
        const int ncol = kNqg * P1;
        const int nseg = (kCluster == 1) ? P.nseg_cta[r] : P.nseg_cl[r];   // host-tuned split of the columns along b
        const int seglen = (P1 + nseg - 1) / nseg;
        const int nitem = ncol * nseg;
        for (int it0 = gtid - lane; it0 < nitem; it0 += kTg) {  // warp-uniform trip count
          const int it = it0 + lane;
          const int seg = it / ncol;
          const int c = it - seg * ncol;
          const int ql = c / P1;
          const int a = c - ql * P1;
          // quadrant: the CTA (cluster rank) owns 8/kCluster octants, an octant has one quadrant per
          // principal axis p (0: z, 1: y, 2: x = branch order of cinterp); octant bit0/1/2 = sign of x/y/z
          const int qc = grp * kNqg + ql;
          const int oct = (int)crank * (8 / kCluster) + qc / 3;
          const int p = qc - (qc / 3) * 3;
          const int sx = (oct & 1) ? -1 : 1, sy = (oct & 2) ? -1 : 1, sz = (oct & 4) ? -1 : 1;
          const int sp = (p == 0) ? sz : (p == 1 ? sy : sx);
          const int sa = (p == 2) ? sy : sx;
          const int sb = (p == 0) ? sy : sz;
          // axes: p==0: (P,A,B)=(z,x,y); p==1: (y,x,z); p==2: (x,y,z)
          const int lrP = (p == 0) ? lr2 : (p == 1 ? lr1 : lr0), llP = (p == 0) ? ll2 : (p == 1 ? ll1 : ll0);
          const int lrA = (p == 2) ? lr1 : lr0, llA = (p == 2) ? ll1 : ll0;
          const int lrB = (p == 0) ? lr1 : lr2, llB = (p == 0) ? ll1 : ll2;
          const bool col_ok = (it < nitem) && r <= (sp > 0 ? lrP : llP) && a <= (sa > 0 ? lrA : llA);
          const int bmax = col_ok ? min(r, sb > 0 ? lrB : llB) : -1;
          const int b0 = seg * seglen;
          const int b1 = min(b0 + seglen - 1, r);      // last b of this segment (loop runs seglen times for all)
          const int nB = (p == 0) ? n1 : n2;
          const int nA = (p == 2) ? n1 : n0;
          const int nP = (p == 0) ? n2 : (p == 1 ? n1 : n0);
          const int srcP = (p == 0) ? src2 : (p == 1 ? src1 : src0);
          const int srcA = (p == 2) ? src1 : src0;
          const int srcB = (p == 0) ? src1 : src2;
          // x-principal quadrants walk planes of constant x: they use the y-fastest twins of the grids
          // (index (x*n2 + z)*n1 + y) so that a warp's lanes (consecutive a = y) stay contiguous in memory
          const unsigned strP = (p == 0) ? st2 : (p == 1 ? st1 : (unsigned)n1 * (unsigned)n2);
          const unsigned strA = 1u;
          const unsigned strideB = (p == 0) ? st1 : (p == 1 ? st2 : (unsigned)n1);
          const double* __restrict__ g_tau = (p == 2) ? P.tau_cell_t : P.tau_cell;
          double* __restrict__ g_phih = (p == 2) ? P.phih_t : P.phih;
          const unsigned base = (unsigned)wrap(srcP + sp * r, nP) * strP + (unsigned)wrap(srcA + sa * a, nA) * strA;
          int posB = srcB + sb * b0;
          if (posB < 0) posB += nB;
          else if (posB >= nB) posB -= nB;
          // the walk along b crosses the periodic boundary at most once: after kwrap more steps
          const int kwrap = (sb > 0) ? (nB - 1 - posB) : posB;
          // column-level geometry
          const double ua = (double)a * inv_r;   // 1-dx of cinterp (a==r gives 1 to an ulp; the cells it would exclude read as 0)
          const double ca2 = (double)(r * r + a * a);
          // ownership pieces that do not depend on b (see header comment)
          const bool own_col = (a > 0 || sa > 0) && (p != 2 || a < r);
          // owned rows of this column: b in [own_lo, own_hi] (b==0 belongs to the sb>0 quadrant; b==r belongs
          // to the z-principal quadrant)
          const int own_lo = own_col ? ((sb > 0) ? 0 : 1) : 0x7fffffff;
          const int own_hi = (p == 0) ? r : r - 1;
          const int loss_b = (sb > 0) ? lrB : llB;     // row on the subbox boundary (any row if the column is)
          const bool loss_col = (sp > 0 ? r == lrP : r == llP) || (sa * a == lrA) || (sa * a == -llA);
          const bool a_in = a <= r - 1;           // column a exists in plane r-1
          const int blast = min(bmax, r - 1);     // last b with an upstream value in this column
          const int nact = max(0, min(b1, bmax) - b0 + 1);   // cells of this segment that exist
          const int nup = max(0, min(b1, blast) - b0 + 1);   // ... that have an upstream cell (a|a-1, b)
          const double* pown = prev + ql * r * r + b0 * r + a;   // plane r-1 patch (stride r), row b0
          double* pout = cur + ql * P1 * P1 + b0 * P1 + a;       // plane r patch (stride r+1)
          // carried upstream values of row b0-1
          double c_own_bm1 = 0.0, c_left_bm1 = 0.0;
          if (b0 >= 1 && b0 - 1 <= blast) {
            if (a_in) c_own_bm1 = pown[-r];
            if (a >= 1) c_left_bm1 = pown[-r - 1];
          }
          // software pipeline: the plane and grid values of the next cell are requested one iteration ahead
          unsigned cell_next = base + (unsigned)posB * strideB;
          double tau_cell_next = 0.0, c_own_next = 0.0, c_lane0_next = 0.0;
          if (nact > 0) tau_cell_next = g_tau[cell_next];
          if (a_in && nup > 0) c_own_next = pown[0];
          if (lane == 0 && a >= 1 && nup > 0) c_lane0_next = pown[-1];
          double bd = (double)b0;
          const int dcell = sb * (int)strideB;
          const int dcell_wrap = dcell - sb * (int)strideB * nB;   // step that crosses the boundary
          for (int k = 0; k < seglen; ++k) {
            const int b = b0 + k;
            const bool active = k < nact;
            // upstream optical depths (cells outside plane r-1 have weight 0; read as 0)
            const double c_own = c_own_next;
            double c_left = __shfl_up_sync(0xffffffffu, c_own, 1);
            if (lane == 0) c_left = c_lane0_next;
            if (a == 0) c_left = 0.0;
            const double t1 = c_left_bm1, t2 = c_own_bm1, t3 = c_left, t4 = c_own;
            c_left_bm1 = c_left;
            c_own_bm1 = c_own;
            const unsigned cell = cell_next;
            const double tau_cell = tau_cell_next;
            // advance to row b+1 and request its inputs
            pown += r;
            cell_next = (unsigned)((int)cell + ((k == kwrap) ? dcell_wrap : dcell));
            c_own_next = 0.0;
            if (k + 1 < nact) tau_cell_next = g_tau[cell_next];
            if (k + 1 < nup) {
              if (a_in) c_own_next = pown[0];
              if (lane == 0 && a >= 1) c_lane0_next = pown[-1];
            } else {
              c_lane0_next = 0.0;
            }
            if (active) {
              bool stop = false;
              // x-fastest index of the cell, only for the grids that have no y-fastest twin
              unsigned xcell = 0;
              if ((kDebug || kLls == 2) && p == 2) {
                const unsigned t = cell;                    // (x*n2 + z)*n1 + y
                const unsigned yy = t % (unsigned)n1, xz = t / (unsigned)n1;
                xcell = (xz % (unsigned)n2) * st2 + yy * st1 + xz / (unsigned)n2;
              }
              // cinterp, column_density.f90:108-171, with a common denominator
              const double ub = bd * inv_r;  // 1-dy
              const double va = 1.0 - ua, vb = 1.0 - ub;
              const double s1 = ua * ub, s2 = ub * va, s3 = ua * vb, s4 = va * vb;
              // weightf = 1/max(0.6, tau), column_density.f90:276-293
              const double m1 = sel_max(t1, 0.6), m2 = sel_max(t2, 0.6);
              const double m3 = sel_max(t3, 0.6), m4 = sel_max(t4, 0.6);
              const double p12 = m1 * m2, p34 = m3 * m4;
              const double e1 = s1 * (m2 * p34), e2 = s2 * (m1 * p34), e3 = s3 * (m4 * p12), e4 = s4 * (m3 * p12);
              const double num = fma(t1, e1, fma(t2, e2, fma(t3, e3, t4 * e4)));
              const double den = (e1 + e2) + (e3 + e4);
              double tau_in = num * fast_rcp(den);
              if (r == 1 && (a == 1 || b == 1)) tau_in *= (a == 1 && b == 1) ? P.sqrt3 : P.sqrt2;  // :152-158
              const double b2 = bd * bd;
              const double q2 = ca2 + b2;
              const double rs = fast_rsqrt(q2);
              const double pathc = q2 * rs * inv_r;                        // sqrt(1+(a^2+b^2)/r^2)
              if (kLls == 3) {  // evolve_point.F90:186-196
                if (dist2_of(P, p, r, a, b2, q2) > P.rmax_lls2) stop = true;
              } else if (kLls == 2) {
                tau_in = fma((double)P.lls_grid[(p == 2) ? xcell : cell] * P.sigma_HI, pathc, tau_in);
              } else if (kLls == 1) {
                tau_in = fma(P.tau_lls, pathc, tau_in);
              }
              if (tau_in > P.tau_stop) stop = true;                          // :201
              const double tau_out = fma(tau_cell, pathc, tau_in);         // :247-248
              *pout = tau_out;
              const bool owner = (b >= own_lo) && (b <= own_hi);
              if (owner) {
                if (kDebug) P.coldens_dbg[(p == 2) ? xcell : cell] = tau_out * P.inv_sigma;
                if (!stop && normflux > 0.0) {
                  double phi_all, phi_out;
                  photo_rates(tau_in, tau_out, normflux, s_thick, s_logtab, P, phi_all, phi_out);
                  // vol_ph = 4*pi*dist2*path (evolve_point.F90:170-177); rate = phi_all/(vol_ph*nHI) = phi_all/volfac
                  const double dist2 = dist2_of(P, p, r, a, b2, q2);
                  const double volfac = P.fourpi_over_sigma * dist2 * pathc * tau_cell;
                  const double inv_vol = fast_rcp(volfac);
                  const double photo_cell = phi_all * inv_vol;             // evolve_point.F90:262
                  if (photo_cell != 0.0) atomicAdd(&g_phih[cell], photo_cell);  // :283-284
                  // boundary of this pass's subbox (:290-295): photo_out*vol/vol_ph
                  if (loss_col || b == loss_b)
                    loss = fma(phi_out * P.vol, inv_vol * (tau_cell * P.inv_sigma_dr0), loss);
                }
              }
            }
            pout += P1;
            bd += 1.0;
          }
        }
        // plane r complete before plane r+1 reads it (quadrants never cross groups)
        if (kGroups == 1) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(kTg) : "memory");
      }
      r_done = rmax;
      // photon_loss_src = sum over the work group (plays photon_loss_src_thread, evolve_source.F90:183-186)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
      if (lane == 0) s_red[tid >> 5] = loss;
      __syncthreads();
      if (kCluster == 1) {
        double t = 0.0;
        for (int i = 0; i < kT / 32; ++i) t += s_red[i];   // same order in every thread
        photon_loss_src = t;
      } else {
        if (tid == 0) {
          double t = 0.0;
          for (int i = 0; i < kT / 32; ++i) t += s_red[i];
          cluster.map_shared_rank(&s_slot[pass_parity][0], 0)[crank] = t;   // DSMEM store into rank 0
        }
        cluster.sync();
        const double* slots = cluster.map_shared_rank(&s_slot[pass_parity][0], 0);
        double t = 0.0;
        for (int i = 0; i < kCluster; ++i) t += slots[i];   // fixed order: identical in all 8 CTAs
        photon_loss_src = t;
        pass_parity ^= 1;
      }
    }
    if (crank == 0 && tid == 0) {
      P.nbox_out[ns] = nbox;              // sum_nbox=sum_nbox+nbox, :219
      P.loss_out[ns] = photon_loss_src;   // photon_loss(1)=photon_loss(1)+photon_loss_src, :216
    }
    if (kCluster > 1) cluster.sync();     // every peer has read s_work before rank 0 fetches the next ticket
  }
  if (kCluster > 1) cluster.sync();  // nobody leaves while a peer may still read rank 0's shared memory
}

// tau_cell = sigma*dr(1) * max(1-max(xh_av,eps),eps) * ndens  (evolve_point.F90:137-145, doric.f90:141-155)
__global__ void taucell_kernel(const float* __restrict__ ndens, const double* __restrict__ xh_av,
                               double* __restrict__ tau_cell, size_t n, double sigma_dr0, double eps) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride) {
    const double h_av1 = fmax(xh_av[c], eps);
    const double h_av0 = fmax(1.0 - h_av1, eps);
    tau_cell[c] = sigma_dr0 * (h_av0 * (double)ndens[c]);
  }
}

__global__ void pair_table_kernel(const double* __restrict__ tab, double2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > kNumTau) return;
  const double lo = tab[i];
  out[i] = make_double2(lo, tab[min(kNumTau, i + 1)] - lo);
}

}  // namespace

size_t raytrace_scratch_doubles_per_cta(int plane_stride) {
  // sized for the single-CTA kernel (24 quadrants); a cluster CTA uses 3 of them
  return (size_t)2 * kQuadrants * plane_stride * plane_stride;
}

static size_t rt_smem_bytes(int plane_doubles) {
  return (size_t)(kTableLen + 128) * sizeof(double2) + (size_t)2 * plane_doubles * sizeof(double);
}

typedef void (*RtKernel)(RtParams);

template <int kT, int kCluster, int kGroups>
static RtKernel pick_kernel(int lls, bool debug) {
  if (debug) {
    switch (lls) {
      case 0: return raytrace_kernel<kT, kCluster, kGroups, 0, true>;
      case 1: return raytrace_kernel<kT, kCluster, kGroups, 1, true>;
      case 2: return raytrace_kernel<kT, kCluster, kGroups, 2, true>;
      default: return raytrace_kernel<kT, kCluster, kGroups, 3, true>;
    }
  }
  switch (lls) {
    case 0: return raytrace_kernel<kT, kCluster, kGroups, 0, false>;
    case 1: return raytrace_kernel<kT, kCluster, kGroups, 1, false>;
    case 2: return raytrace_kernel<kT, kCluster, kGroups, 2, false>;
    default: return raytrace_kernel<kT, kCluster, kGroups, 3, false>;
  }
}

// work-group shapes of the many-CTA kernel (C2B_CLUSTER_VARIANT selects one)
struct ClusterVariant {
  int threads, cluster, groups;
  RtKernel (*pick)(int, bool);
  int ctas_per_sm() const { return threads <= 256 ? 2 : 1; }
};
static const ClusterVariant kVariants[] = {
    {512, 8, 1, pick_kernel<512, 8, 1>},  // 0: one octant per CTA, one CTA per SM (planes in shared memory up to r = 63)
    {480, 8, 3, pick_kernel<480, 8, 3>},  // 1: as 0 with one warp group (own named barrier) per face quadrant
    {256, 8, 1, pick_kernel<256, 8, 1>},  // 2: one octant per CTA, two CTAs (two sources) per SM  [default]
    {256, 2, 1, pick_kernel<256, 2, 1>},  // 3: four octants per CTA, two CTAs per SM
};
static int g_variant = 2;

static void cluster_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, const ClusterVariant& v, int nclusters,
                           size_t smem, cudaStream_t stream) {
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3(v.cluster * nclusters, 1, 1);
  cfg->blockDim = dim3(v.threads, 1, 1);
  cfg->dynamicSmemBytes = smem;
  cfg->stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = v.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
}

int raytrace_configure(int max_radius, RtLaunchInfo* info) {
  int dev = 0, max_optin = 0, sm_total = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&sm_total, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (const char* env = getenv("C2B_CLUSTER_VARIANT")) {
    const int v = atoi(env);
    if (v >= 0 && v < (int)(sizeof(kVariants) / sizeof(kVariants[0]))) g_variant = v;
  }
  const ClusterVariant& V = kVariants[g_variant];
  const size_t fixed = (size_t)(kTableLen + 128) * sizeof(double2);
  // ---- single-CTA kernel: two CTAs per SM share the opt-in shared memory ----------------------
  const int per_cta = std::min(max_optin, sm_total / kCtaPerSm - 2048);
  int cap = (int)(((size_t)per_cta - fixed - 1024) / (2 * sizeof(double)));
  cap = std::min(cap, kQuadrants * (max_radius + 1) * (max_radius + 1)) & ~1;
  // ---- cluster kernel: one CTA per SM with all of the opt-in shared memory ----------------------
  const int per_cta_cl = (V.ctas_per_sm() == 2) ? std::min(max_optin, sm_total / 2 - 2048) : max_optin;
  int cap_cl = (int)(((size_t)per_cta_cl - fixed - 2048) / (2 * sizeof(double)));
  cap_cl = std::min(cap_cl, (kQuadrants / V.cluster) * (max_radius + 1) * (max_radius + 1) + 2 * V.groups) & ~1;
  for (int dbg = 0; dbg < 2; ++dbg)
    for (int lls = 0; lls < 4; ++lls) {
      cudaError_t e = cudaFuncSetAttribute(pick_kernel<kThreadsCta, 1, 1>(lls, dbg != 0),
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rt_smem_bytes(cap));
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(V.pick(lls, dbg != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)rt_smem_bytes(cap_cl));
      if (e != cudaSuccess) return (int)e;
    }
  info->smem_plane_doubles = cap;
  info->smem_plane_doubles_cl = cap_cl;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pick_kernel<kThreadsCta, 1, 1>(1, false), kThreadsCta,
                                                rt_smem_bytes(cap));
  info->grid_cta = sms * std::max(1, per_sm);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  cluster_config(&cfg, attr, V, sms, rt_smem_bytes(cap_cl), nullptr);
  int nclusters = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, V.pick(1, false), &cfg);
  if (e != cudaSuccess) return (int)e;
  info->clusters = std::max(1, nclusters);
  info->cluster_size = V.cluster;
  // scratch slots of 2*24*S*S doubles: one per resident CTA of the single-CTA kernel plus one per resident
  // cluster (its CTAs share a slot: 24/cluster quadrants each)
  info->grid_max = info->grid_cta + info->clusters;
  return 0;
}

// Number of b-segments per column for shell r: minimises rounds x (segment length + set-up) for a group
// of `threads` threads that owns `nq` quadrants, i.e. nq*(r+1) columns.
static void build_nseg_table(int max_radius, int nq, int threads, std::vector<int>& tab) {
  tab.assign((size_t)max_radius + 2, 1);
  for (int r = 1; r <= max_radius; ++r) {
    const int P1 = r + 1, ncol = nq * P1;
    long best = -1;
    int best_n = 1;
    for (int n = 1; n <= std::max(1, P1 / 2) && n <= 64; ++n) {
      const long rounds = ((long)ncol * n + threads - 1) / threads;
      const long cost = rounds * ((P1 + n - 1) / n + 2);
      if (best < 0 || cost < best) { best = cost; best_n = n; }
    }
    tab[r] = best_n;
  }
}

void raytrace_nseg_tables(int max_radius, std::vector<int>& cta, std::vector<int>& cl) {
  const ClusterVariant& V = kVariants[g_variant];
  build_nseg_table(max_radius, kQuadrants, kThreadsCta, cta);
  build_nseg_table(max_radius, kQuadrants / V.cluster / V.groups, V.threads / V.groups, cl);
}

static int lls_mode(const RtParams& p) { return p.use_lls ? p.type_lls : 0; }

void launch_raytrace(const RtParams& p, int grid, cudaStream_t stream) {
  pick_kernel<kThreadsCta, 1, 1>(lls_mode(p), p.coldens_dbg != nullptr)
      <<<grid, kThreadsCta, rt_smem_bytes(p.smem_plane_doubles), stream>>>(p);
}

int launch_raytrace_cluster(const RtParams& p, int nclusters, cudaStream_t stream) {
  const ClusterVariant& V = kVariants[g_variant];
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  cluster_config(&cfg, attr, V, nclusters, rt_smem_bytes(p.smem_plane_doubles_cl), stream);
  return (int)cudaLaunchKernelEx(&cfg, V.pick(lls_mode(p), p.coldens_dbg != nullptr), p);
}

void launch_taucell(const float* ndens, const double* xh_av, double* tau_cell, size_t n, double sigma_dr0,
                    double eps, cudaStream_t stream) {
  taucell_kernel<<<chemistry_blocks(), 256, 0, stream>>>(ndens, xh_av, tau_cell, n, sigma_dr0, eps);
}

void launch_pair_table(const double* tab, double2* out, cudaStream_t stream) {
  pair_table_kernel<<<(kTableLen + 127) / 128, 128, 0, stream>>>(tab, out);
}

}  // namespace c2b
