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
// shell r-1 only.  A shell is the surface of a cube: 6 faces (principal axis p = the dominant
// |offset|, which selects the cinterp branch, and the sign of the principal offset), each split into
// 4 quadrants by the signs of the two transverse offsets.  In quadrant-local coordinates (a,b) =
// transverse distances, the cell (a,b) of plane r reads (a-1|a, b-1|b) of plane r-1 of the SAME
// quadrant.  Cells shared between quadrants (on-axis a==0 / b==0, cube edges a==r / b==r) are
// computed by every quadrant that needs them as an upstream value (the three cinterp branches agree
// on ties), but only the quadrant the reference's branch order selects "owns" the cell: it alone
// adds the rate into phih_grid and counts the boundary loss.
//
// Mapping.  One work group per source, persistent work groups pulling sources from an atomic ticket
// (the device-side do_grid_master, master_slave.F90:124-231).  A work group is one CTA of 256
// threads (two resident per SM) or, when there are too few long traces to fill the GPU that way, a
// thread-block cluster of 6 CTAs, one per cube face, whose boundary-loss partial sums meet in rank
// 0's shared memory over DSMEM once per subbox pass, or -- for the first subbox of the many short
// traces of early reionization -- a single warp (raytrace_warp_kernel below; what outgrows the
// first subbox is handed over to the one-CTA shape on the device).  Within a shell a thread owns a column (face,
// transverse index a), or a b-segment of it, and walks b, updating the FOUR quadrants of its face
// together: the interpolation weights, the path length and the dilution volume depend only on
// (r,a,b), so they are computed once per four cells, and the four dependency chains interleave in
// the FP64 pipe.  The two upstream values of row b-1 stay in registers from one b to the next; row
// b comes from the plane of shell r-1, which stores the four quadrant values of (a,b) contiguously
// (32 bytes: two 128-bit loads fetch them).  The planes of shell r-1 and r live in shared memory
// while they fit, afterwards in a per-CTA global scratch.  The optical-depth table is staged in
// shared memory as (value, forward difference) pairs.  Faces whose principal axis is x walk planes
// of constant x; they read and accumulate into y-fastest twins of the grids so that their accesses
// are contiguous too (when the pass has enough work to pay for the transposes: RtParams::use_twins).
//
// Upstream cells that do not exist in plane r-1 (a-1 < 0, b-1 < 0, a == r, b == r) are read as
// whatever finite value the buffer holds: their bilinear weight is exactly 0 (the weights are built
// so that a==r / b==r give exactly 1 and 0), so they do not contribute.  The plane buffers are
// zero-filled once (shared memory at kernel start, the global scratch at allocation) and only ever
// receive finite optical depths.
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

#ifndef C2B_RT_THREADS
#define C2B_RT_THREADS 256
#endif
#ifndef C2B_CTA_PER_SM
#define C2B_CTA_PER_SM 2
#endif
constexpr int kT = C2B_RT_THREADS; // threads per CTA
constexpr int kFaces = 6;
constexpr int kClusterSize = 6;    // the many-CTA work group: one CTA per face
constexpr int kCtaPerSm = C2B_CTA_PER_SM;
#ifndef C2B_RT_ILP
#define C2B_RT_ILP 4
#endif
constexpr int kIlp = C2B_RT_ILP;   // quadrants whose interpolation chains are interleaved (4, 2 or 1)
// The six faces of a source never exchange data, so a CTA walks them in independent groups of kFpg faces, each
// group with kT/(6/kFpg) threads and its own barrier (a warp barrier when the group is one warp).
#ifndef C2B_RT_FPG
#define C2B_RT_FPG 6
#endif
constexpr int kFpgCta = C2B_RT_FPG;                      // faces per group, single-CTA kernel
constexpr int kTgCta = kT / (kFaces / kFpgCta);          // threads per group
static_assert(kFaces % kFpgCta == 0 && kT % (kFaces / kFpgCta) == 0 && kTgCta % 32 == 0, "bad face grouping");
constexpr int kPadFront = 8;       // doubles in front of every plane buffer (the a-1 read of column 0)

// per-face constants of the current subbox pass, rebuilt in shared memory once per pass
struct Face {
  int srcP, srcA, srcB;       // 0-based source position along the principal / a / b axis
  int nP, nA, nB;             // mesh extents along them
  int sp;                     // sign of the principal offset
  int p;                      // 0: z, 1: y, 2: x principal (the branch order of cinterp)
  unsigned strP, strB;        // element strides of the principal and the b axis
  int limP;                   // extent of this pass's subbox along sp*P
  int lrA, llA, lrB, llB;     // ... along +a, -a, +b, -b
  unsigned strA;              // stride of the a axis: 1 (contiguous) except for the x-principal faces of the per-warp
                              // kernel, which work on the x-fastest grids directly
  int lay;                    // 2: this face addresses the y-fastest twins, 0: the x-fastest grids
  double dP2, dA2, dB2;       // dr^2 per axis (dist2 of evolve_point.F90:170-174)
  const double* tau;          // opacity grid (x-fastest, or the y-fastest twin for p == 2)
  double* phih;               // rate grid, same layout
  double* heat;               // heating-rate grid (phiheat_grid or its twin), non-isothermal only
};

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

// constants of the table look-up kept in registers / uniform registers for the whole kernel
struct LogC {
  double c0, c1, c2, c3, c4, B;
};

// table coordinate odpos = 1 + (log10(max(1e-20,tau)) - minlogtau)/dlogtau of
// set_tau_table_positions (radiation_photoionrates.F90:184-208), clamped to NumTau.
// logtab[j] = {1/c_j, A + B*log2(c_j)}, c_j = 1 + (j+0.5)/128; coef = B/ln2 * {1,-1/2,1/3,-1/4,1/5}
__device__ __forceinline__ double table_coord(double tau, const double2* __restrict__ logtab, const LogC& L) {
  const double t = sel_max(tau, 1.0e-20);
  const int hi = __double2hiint(t);
  const int lo = __double2loint(t);
  const int j = (hi >> 13) & 127;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const double2 lt = logtab[j];
  // (double)(exponent) without an integer->double conversion: 2^52 + (e + 2^31) as bits, minus 2^52 + 2^31
  const double ed = __hiloint2double(0x43300000, ((hi >> 20) - 1023) ^ 0x80000000) - 4503601774854144.0;
  const double rr = fma(m, lt.x, -1.0);
  double q = fma(rr, L.c4, L.c3);
  q = fma(rr, q, L.c2);
  q = fma(rr, q, L.c1);
  q = fma(rr, q, L.c0);
  double od = fma(L.B, ed, lt.y);
  od = fma(rr, q, od);
  return pos_min(od, (double)kNumTau);  // od >= 0 because tau >= 1e-20
}

// read_table (radiation_photoionrates.F90:212-228) on the (value, forward difference) pairs.
// ipos = int(odpos), residual = odpos - ipos, by magic-number rounding of odpos-0.5 (no F2I/I2F);
// at an exact integer odpos the pair (ipos-1, 1.0) may come out instead of (ipos, 0.0): same value.
__device__ __forceinline__ double lerp_pairs(const double2* __restrict__ tab, double od, int& ipos, double& res) {
  const double sh = (od - 0.5) + 6755399441055744.0;
  const int ip = __double2loint(sh);    // 0 <= od <= NumTau, so 0 <= ip <= NumTau
  const double fl = sh - 6755399441055744.0;
  ipos = ip;
  res = od - fl;
  const double2 t = tab[ip];
  return fma(t.y, res, t.x);
}

// photoion_rates / photo_lookuptable (radiation_photoionrates.F90:71-317) for one stellar source:
// Gamma_cell*vol_ph = F*(thick(tau_in)-thick(tau_out)), or F*dtau*thin(tau_in) below tau_photo_limit.
// kHeat adds heat_lookuptable (:323-417) on the same table positions: heat*vol_ph = F*(H(tau_in)-H(tau_out)), or
// F*dtau*Hthin(tau_in) below tau_heat_limit.
template <bool kHeat>
__device__ __forceinline__ void photo_rates(double tau_in, double tau_out, double normflux,
                                            const double2* __restrict__ s_thick, const double2* __restrict__ s_logtab,
                                            const double2* __restrict__ s_heat, const RtParams& P, const LogC& L,
                                            double& phi_all, double& phi_out, double& heat_all) {
  // both table coordinates are formed together (two independent chains); the thin branch overrides
  const double od_in = table_coord(tau_in, s_logtab, L);
  const double od_out = table_coord(tau_out, s_logtab, L);
  int ipos, ipos2;
  double res, res2;
  const double phi_in = normflux * lerp_pairs(s_thick, od_in, ipos, res);
  phi_out = normflux * lerp_pairs(s_thick, od_out, ipos2, res2);
  phi_all = phi_in - phi_out;
  const double dtau = tau_out - tau_in;
  if (!(fabs(dtau) > P.tau_photo_limit)) {
    const double lo = P.thin[ipos];
    const double th = lo + (P.thin[min(kNumTau, ipos + 1)] - lo) * res;
    phi_all = normflux * dtau * th;
    phi_out = phi_in - phi_all;
  }
  if (kHeat) {
    const double2 hi_ = s_heat[ipos], ho_ = s_heat[ipos2];
    heat_all = normflux * fma(hi_.y, res, hi_.x) - normflux * fma(ho_.y, res2, ho_.x);
    if (!(fabs(dtau) > P.tau_heat_limit)) {
      const double lo = P.heat_thin[ipos];
      heat_all = normflux * dtau * (lo + (P.heat_thin[min(kNumTau, ipos + 1)] - lo) * res);
    }
  }
}

// x-fastest index of a cell addressed in the layout of face p (p == 2: the y-fastest twin, (x*n2 + z)*n1 + y);
// only for the grids that have no y-fastest twin (LLS_grid, the coldensh_out diagnostic)
__device__ __forceinline__ unsigned xfast_index(const RtParams& P, int p, unsigned cell) {
  if (p != 2) return cell;
  const unsigned n1 = (unsigned)P.n[1], n2 = (unsigned)P.n[2];
  const unsigned yy = cell % n1, xz = cell / n1;
  return (xz % n2) * ((unsigned)P.n[0] * n1) + yy * (unsigned)P.n[0] + xz / n2;
}

__device__ __forceinline__ int wrap(int x, int n) {
  // modulo(x-1,mesh)+1 of evolve_point.F90:122-124 for |offset| <= n/2, 0-based
  if (x < 0) x += n;
  else if (x >= n) x -= n;
  return x;
}

// synchronises the kTg threads of face group g
template <int kTg>
__device__ __forceinline__ void group_sync(int g) {
  if (kTg == 32) __syncwarp();
  else if (kTg == kT) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(kTg) : "memory");
}

struct Quad {
  double v[4];
};
__device__ __forceinline__ Quad load_quad(const double* p) {
  const double2 lo = *reinterpret_cast<const double2*>(p);
  const double2 hi = *reinterpret_cast<const double2*>(p + 2);
  Quad q;
  q.v[0] = lo.x; q.v[1] = lo.y; q.v[2] = hi.x; q.v[3] = hi.y;
  return q;
}
__device__ __forceinline__ void store_quad(double* p, const Quad& q) {
  *reinterpret_cast<double2*>(p) = make_double2(q.v[0], q.v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(q.v[2], q.v[3]);
}

// per-source values every thread of the work group holds
struct SrcCtx {
  double normflux;
  int reach;      // subboxsize*nbox of this pass
  int rsafe;      // shells r < rsafe are not clipped by the periodic half box on any side
  bool emit;      // this pass adds rates and counts the boundary loss (false: a dark source, or a pass that only
                  // rebuilds the planes of shells the per-warp kernel has already emitted)
};

// One shell of one source: the cells with Chebyshev distance r of the faces this CTA owns.
// kGlobal: planes in the global scratch (else shared memory); kClip: the shell may touch the limits of
// the subbox / half box; kR1: r == 1 (the sqrt(2)/sqrt(3) factors of column_density.f90:152-158).
// kLls: 1 = homogeneous (tau_lls may be 0: no LLS), 2 = LLS_grid, 3 = R_max barrier (LLS.F90:107-116).
// The kFpg faces `faces[0..kFpg)` are walked by kTg threads (this thread is number `gtid`); prev/cur are the plane
// buffers of the first face, those of the next faces follow at `fstride` doubles.  min_hi returns the high word of
// the smallest optical depth written (the dead-face rule in the kernel).
template <int kFpg, int kTg, bool kGlobal, bool kClip, bool kR1, int kLls, bool kDebug, bool kHeat>
__device__ __forceinline__ void trace_shell(const RtParams& P, const SrcCtx& S, const Face* __restrict__ faces, int gtid,
                                            const double2* __restrict__ s_thick, const double2* __restrict__ s_logtab,
                                            const double2* __restrict__ s_heat, const LogC& L, int r, const double* __restrict__ prev,
                                            double* __restrict__ cur, int fstride, int nseg, int* next_item,
                                            double& loss, int& min_hi) {
  if (kGlobal) {
    __builtin_assume(__isGlobal(prev));
    __builtin_assume(__isGlobal(cur));
  } else {
    __builtin_assume(__isShared(prev));
    __builtin_assume(__isShared(cur));
  }
  const int P1 = r + 1;
  const int seglen = (P1 + nseg - 1) / nseg;
  const int ncol = kFpg * P1;
  const int nitem = ncol * nseg;
  const float inv_ncol = 1.0f / (float)ncol, inv_P1 = 1.0f / (float)P1;
  const double rd = (double)r;
  const double inv_r = fast_rcp(rd);
  const double r2d = rd * rd;
  const bool loss_shell = (r == S.reach);   // non-clipped shells: every cell of the shell is on the subbox boundary, or none
  // The work items of the shell are handed out 32 at a time from a counter in shared memory: a warp whose cells
  // are cheap (stopped behind the 2e19 column: no rates) simply takes more of them, so the warps of the group
  // reach the barrier of the shell together.
  for (;;) {
    int it = 0;
    if ((threadIdx.x & 31) == 0) it = atomicAdd(next_item, 32);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= nitem) break;
    it += (int)(threadIdx.x & 31);
    if (it < nitem) do {   // (a `continue` below leaves this item)
    // work item = (segment of b, face, column a), a fastest so that a warp spans adjacent columns
    const int seg = (nseg == 1) ? 0 : (int)(((float)it + 0.5f) * inv_ncol);
    const int c = it - seg * ncol;
    const int flg = (kFpg == 1) ? 0 : (int)(((float)c + 0.5f) * inv_P1);
    const int a = c - flg * P1;
    const Face& F = faces[flg];
    const int b0 = seg * seglen;
    int bend = min(b0 + seglen - 1, r);
    // ownership (see header): columns (a > 0 || sa > 0) && (p != 2 || a < r); bit j = quadrant (sa<0) | (sb<0)<<1
    unsigned colmask = (a > 0) ? 0xFu : 0x5u;
    if (F.p == 2 && a == r) colmask = 0u;
    unsigned losscol = 0u;
    if (kClip) {
      if (r > F.limP) continue;                       // the face lies outside this pass's subbox
      if (a > max(F.lrA, F.llA)) continue;
      bend = min(bend, max(F.lrB, F.llB));
      if (a > F.lrA) colmask &= ~0x5u;
      if (a > F.llA) colmask &= ~0xAu;
      // cells on the boundary of this pass's subbox (evolve_point.F90:290-295)
      if (r == F.limP) losscol = 0xFu;
      if (a == F.lrA) losscol |= 0x5u;
      if (a == F.llA) losscol |= 0xAu;
    }
    if (!S.emit) colmask = 0u;                        // a dark source is never traced (evolve_source.F90:119-131)
    const int nrow = bend - b0 + 1;
    if (nrow <= 0) continue;
    const int own_hi = (F.p == 0) ? r : r - 1;        // b == r belongs to the z-principal face
    // addresses: cell = posP*strP + posA + posB*strB
    const unsigned rowP = (unsigned)wrap(F.srcP + F.sp * r, F.nP) * F.strP;
    const unsigned baseAp = rowP + (unsigned)wrap(F.srcA + a, F.nA) * F.strA;
    const unsigned baseAm = rowP + (unsigned)wrap(F.srcA - a, F.nA) * F.strA;
    const int nB = F.nB;
    const unsigned strB = F.strB;
    int posBp = wrap(F.srcB + b0, nB), posBm = wrap(F.srcB - b0, nB);
    const double* __restrict__ g_tau = F.tau;
    double* __restrict__ g_phih = F.phih;
    double* __restrict__ g_heat = F.heat;
    __builtin_assume(__isGlobal(g_tau));
    __builtin_assume(__isGlobal(g_phih));
    if (kHeat) __builtin_assume(__isGlobal(g_heat));
    // column geometry shared by the four quadrants
    const double ad = (double)a;
    const double ua = (a == r) ? 1.0 : ad * inv_r;    // 1-dx of cinterp
    const double va = 1.0 - ua;
    const double a2d = ad * ad;
    const double ca2 = r2d + a2d;
    const double cA = fma(F.dP2, r2d, F.dA2 * a2d);
    const double dB2 = F.dB2;
    // planes of this face: [b][a][quadrant]
    const double* pp = prev + flg * fstride + ((b0 * r + a) << 2);
    double* pc = cur + flg * fstride + ((b0 * P1 + a) << 2);
    const int pstep = r << 2, cstep = P1 << 2;
    Quad t1, t2;   // upstream values of row b-1: (a-1,b-1), (a,b-1)
    if (b0 >= 1) {
      t1 = load_quad(pp - pstep - 4);
      t2 = load_quad(pp - pstep);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) t1.v[j] = t2.v[j] = 0.0;
    }
    // software pipeline: the plane and grid values of the next row are requested one iteration ahead
    Quad own_next = load_quad(pp);
    unsigned offp = (unsigned)posBp * strB, offm = (unsigned)posBm * strB;
    Quad tc_next;
    tc_next.v[0] = g_tau[baseAp + offp];
    tc_next.v[1] = g_tau[baseAm + offp];
    tc_next.v[2] = g_tau[baseAp + offm];
    tc_next.v[3] = g_tau[baseAm + offm];
    double bd = (double)b0;
    for (int k = 0; k < nrow; ++k) {
      const int b = b0 + k;
      const Quad t4 = own_next;               // (a, b)
      const Quad t3 = load_quad(pp - 4);      // (a-1, b)
      const Quad tc = tc_next;
      const unsigned cellp = offp, cellm = offm;
      // advance to row b+1 and request its inputs
      pp += pstep;
      posBp += 1;
      offp += strB;
      if (posBp == nB) { posBp = 0; offp = 0u; }
      posBm -= 1;
      offm -= strB;
      if (posBm < 0) { posBm = nB - 1; offm = (unsigned)(nB - 1) * strB; }
      if (k + 1 < nrow) {
        own_next = load_quad(pp);
        tc_next.v[0] = g_tau[baseAp + offp];
        tc_next.v[1] = g_tau[baseAm + offp];
        tc_next.v[2] = g_tau[baseAp + offm];
        tc_next.v[3] = g_tau[baseAm + offm];
      }
      // ---- geometry of (r,a,b), common to the four quadrants --------------------------------------
      const double ub = (b == r) ? 1.0 : bd * inv_r;  // 1-dy
      const double vb = 1.0 - ub;
      const double s1 = ua * ub, s2 = ub * va, s3 = ua * vb, s4 = va * vb;   // column_density.f90:137-149
      const double b2 = bd * bd;
      const double q2 = ca2 + b2;
      const double rs = fast_rsqrt(q2);
      const double pathc = q2 * rs * inv_r;                    // sqrt(1+(a^2+b^2)/r^2)
      const double dist2 = fma(dB2, b2, cA);                   // evolve_point.F90:170-174
      const double volk = P.fourpi_over_sigma * dist2 * pathc; // vol_ph*nHI = volk*tau_cell (:176-177)
      bool stop_all = false;
      if (kLls == 3) stop_all = dist2 > P.rmax_lls2;           // evolve_point.F90:186-196
      double corr = 1.0;
      if (kR1) corr = (a == 1 && b == 1) ? P.sqrt3 : ((a == 1 || b == 1) ? P.sqrt2 : 1.0);   // column_density.f90:152-158
      unsigned rowmask = (b <= own_hi) ? ((b > 0) ? 0xFu : 0x3u) : 0u;
      unsigned lossmask = losscol;
      if (kClip) {
        if (b > F.lrB) rowmask &= ~0x3u;
        if (b > F.llB) rowmask &= ~0xCu;
        if (b == F.lrB) lossmask |= 0x3u;
        if (b == F.llB) lossmask |= 0xCu;
      } else if (loss_shell) {
        lossmask = 0xFu;
      }
      const unsigned ownmask = colmask & rowmask;
      // The quadrants are processed kIlp at a time: phase A (interpolation, branch-free so that the kIlp chains
      // interleave in the FP64 pipe), then phase B (rates of the owned, unstopped cells) for the same quadrants.
      Quad out;
#pragma unroll
      for (int j0 = 0; j0 < 4; j0 += kIlp) {
        double tin[kIlp];
#pragma unroll
        for (int jj = 0; jj < kIlp; ++jj) {
          const int j = j0 + jj;
          // cinterp, column_density.f90:108-171, with a common denominator; weightf = 1/max(0.6, tau), :276-293
          const double m1 = sel_max(t1.v[j], 0.6), m2 = sel_max(t2.v[j], 0.6);
          const double m3 = sel_max(t3.v[j], 0.6), m4 = sel_max(t4.v[j], 0.6);
          const double p12 = m1 * m2, p34 = m3 * m4;
          const double e1 = s1 * (m2 * p34), e2 = s2 * (m1 * p34), e3 = s3 * (m4 * p12), e4 = s4 * (m3 * p12);
          const double num = fma(t1.v[j], e1, fma(t2.v[j], e2, fma(t3.v[j], e3, t4.v[j] * e4)));
          const double den = (e1 + e2) + (e3 + e4);
          double tau_in = num * fast_rcp(den);
          if (kR1) tau_in *= corr;
          if (kLls == 2) {
            const unsigned cell = ((j & 1) ? baseAm : baseAp) + ((j & 2) ? cellm : cellp);
            tau_in = fma((double)P.lls_grid[xfast_index(P, F.lay, cell)] * P.sigma_HI, pathc, tau_in);
          } else if (kLls == 1) {
            tau_in = fma(P.tau_lls, pathc, tau_in);
          }
          tin[jj] = tau_in;
          out.v[j] = fma(tc.v[j], pathc, tau_in);       // evolve_point.F90:247-248
          min_hi = min(min_hi, __double2hiint(out.v[j]));
        }
#pragma unroll
        for (int jj = 0; jj < kIlp; ++jj) {
          const int j = j0 + jj;
          if ((ownmask >> j) & 1u) {
            const unsigned cell = ((j & 1) ? baseAm : baseAp) + ((j & 2) ? cellm : cellp);
            if (kDebug) P.coldens_dbg[xfast_index(P, F.lay, cell)] = out.v[j] * P.inv_sigma;
            if (!(tin[jj] > P.tau_stop) && !stop_all) {    // evolve_point.F90:201
              const double tau_cell = tc.v[j];
              double phi_all, phi_out, heat_all = 0.0;
              photo_rates<kHeat>(tin[jj], out.v[j], S.normflux, s_thick, s_logtab, s_heat, P, L, phi_all, phi_out, heat_all);
              // vol_ph = 4*pi*dist2*path (evolve_point.F90:170-177); rate = phi_all/(vol_ph*nHI)
              const double inv_vol = fast_rcp(volk * tau_cell);
              const double photo_cell = phi_all * inv_vol;            // :262
              if (photo_cell != 0.0) atomicAdd(&g_phih[cell], photo_cell);   // :283-284
              if (kHeat) {   // phi%heat = heat_all/vol_ph, not divided by the neutral density (:285-286)
                const double heat_cell = heat_all * (inv_vol * (tau_cell * P.inv_sigma_dr0));
                if (heat_cell != 0.0) atomicAdd(&g_heat[cell], heat_cell);
              }
              // boundary of this pass's subbox (:290-295): photo_out*vol/vol_ph
              if ((lossmask >> j) & 1u) loss = fma(phi_out * P.vol, inv_vol * (tau_cell * P.inv_sigma_dr0), loss);
            }
          }
        }
      }
      store_quad(pc, out);
      pc += cstep;
      t1 = t3;
      t2 = t4;
      bd += 1.0;
    }
    } while (0);
  }
}

// per-face constants of one subbox pass.  face f: p = f mod 3 (0: z, 1: y, 2: x principal), positive side first.
// axes: p==0: (P,A,B)=(z,x,y); p==1: (y,x,z); p==2: (x,y,z) on the y-fastest twins (index (x*n2+z)*n1+y)
// kTwins = false: the x-principal faces address the x-fastest grids too (a = y has stride n0, b = z has n0*n1)
template <bool kTwins>
__device__ __forceinline__ Face make_face(const RtParams& P, int f, int src0, int src1, int src2, int lr0, int ll0,
                                          int lr1, int ll1, int lr2, int ll2) {
  const int p = (f >= 3) ? f - 3 : f, sp = (f >= 3) ? -1 : 1;
  const int n0 = P.n[0], n1 = P.n[1], n2 = P.n[2];
  Face F;
  F.p = p; F.sp = sp; F.strA = 1u; F.lay = (kTwins && p == 2) ? 2 : 0;
  F.srcP = (p == 0) ? src2 : (p == 1 ? src1 : src0);
  F.srcA = (p == 2) ? src1 : src0;
  F.srcB = (p == 0) ? src1 : src2;
  F.nP = (p == 0) ? n2 : (p == 1 ? n1 : n0);
  F.nA = (p == 2) ? n1 : n0;
  F.nB = (p == 0) ? n1 : n2;
  F.strP = (p == 0) ? (unsigned)n0 * (unsigned)n1 : (p == 1 ? (unsigned)n0 : (unsigned)n1 * (unsigned)n2);
  F.strB = (p == 0) ? (unsigned)n0 : (p == 1 ? (unsigned)n0 * (unsigned)n1 : (unsigned)n1);
  const int lrP = (p == 0) ? lr2 : (p == 1 ? lr1 : lr0), llP = (p == 0) ? ll2 : (p == 1 ? ll1 : ll0);
  F.limP = (sp > 0) ? lrP : llP;
  F.lrA = (p == 2) ? lr1 : lr0; F.llA = (p == 2) ? ll1 : ll0;
  F.lrB = (p == 0) ? lr1 : lr2; F.llB = (p == 0) ? ll1 : ll2;
  F.dP2 = (p == 0) ? P.dr2[2] : (p == 1 ? P.dr2[1] : P.dr2[0]);
  F.dA2 = (p == 2) ? P.dr2[1] : P.dr2[0];
  F.dB2 = (p == 0) ? P.dr2[1] : P.dr2[2];
  F.tau = (p == 2) ? P.tau_cell_t : P.tau_cell;
  F.phih = (p == 2) ? P.phih_t : P.phih;
  F.heat = (p == 2) ? P.phiheat_t : P.phiheat;
  if (!kTwins && p == 2) {
    F.strP = 1u; F.strA = (unsigned)n0; F.strB = (unsigned)n0 * (unsigned)n1;
    F.tau = P.tau_cell; F.phih = P.phih; F.heat = P.phiheat;
  }
  return F;
}

// One CTA (kCluster == 1, all six faces) or one cluster of 6 CTAs (one face each) per source.
template <int kCluster, int kLls, bool kDebug, bool kHeat>
__global__ void __launch_bounds__(kT, kCtaPerSm) raytrace_kernel(RtParams P) {
  constexpr int kNf = kFaces / kCluster;            // faces handled by this CTA
  extern __shared__ double2 smem2[];
  double2* s_thick = smem2;                         // kTableLen pairs
  double2* s_logtab = smem2 + kTableLen;            // 128 pairs
  double2* s_heat = smem2 + kTableLen + 128;        // kTableLen pairs (non-isothermal only)
  double* s_planes = reinterpret_cast<double*>(smem2 + kTableLen + 128 + (kHeat ? kTableLen : 0));
  __shared__ Face s_face[kNf];
  __shared__ double s_red[kT / 32];
  __shared__ double s_slot[2][8];                   // per-face boundary loss, alternating by pass (rank 0's copy is used)
  __shared__ int s_work;

  constexpr int kFpg = (kCluster == 1) ? kFpgCta : 1;   // faces per group
  constexpr int kTg = (kCluster == 1) ? kTgCta : kT;    // threads per group
  constexpr int kGroups = kNf / kFpg;
  __shared__ int s_gmin[kGroups][3];                // dead-face rule: smallest high word per group, rotating by shell
  __shared__ int s_next[kGroups][2];                // next work item of the shell, alternating by shell parity
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int g = tid / kTg;                          // face group of this thread
  const int gtid = tid - g * kTg;
  const int fl = g * kFpg;                          // first face (local to this CTA) of the group
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = (kCluster > 1) ? cluster.block_rank() : 0u;
  const int cap = (kCluster == 1) ? P.smem_plane_doubles : P.smem_plane_doubles_cl;   // one shared plane buffer of one face
  for (int i = tid; i < kTableLen; i += kT) s_thick[i] = P.thick2[i];
  for (int i = tid; i < 128; i += kT) s_logtab[i] = P.logtab[i];
  if (kHeat)
    for (int i = tid; i < kTableLen; i += kT) s_heat[i] = P.heat2[i];
  for (int i = tid; i < 2 * kNf * (cap + kPadFront); i += kT) s_planes[i] = 0.0;
  // every face owns two fixed plane buffers (faces run decoupled, so their storage must not move with r):
  // `cap` doubles each in shared memory, Gf each in the global scratch slot of the work group
  double* sbuf0 = s_planes + (size_t)(2 * fl) * (cap + kPadFront) + kPadFront;
  double* sbuf1 = sbuf0 + cap + kPadFront;
  const size_t Gf = raytrace_face_doubles(P.plane_stride);
  const size_t slot = (kCluster == 1) ? (size_t)blockIdx.x : (size_t)(blockIdx.x / kCluster);
  double* gbuf0 = P.scratch + (slot * 2 * kFaces + 2 * ((size_t)crank * kNf + fl)) * Gf + kPadFront;
  double* gbuf1 = gbuf0 + Gf;
  const int sstride = 2 * (cap + kPadFront), gstride = (int)(2 * Gf);   // from a face's buffers to the next face's
  LogC L;
  L.c0 = P.logc[0]; L.c1 = P.logc[1]; L.c2 = P.logc[2]; L.c3 = P.logc[3]; L.c4 = P.logc[4]; L.B = P.logB;
  SrcCtx S;
  S.rsafe = min(min(min(P.lim[0][0], P.lim[0][1]), min(P.lim[1][0], P.lim[1][1])), min(P.lim[2][0], P.lim[2][1]));
  int pass_parity = 0;
  // high word of a column safely above sigma*2e19: a plane whose every high word exceeds it is dead
  const int dead_hi = __double2hiint(P.tau_stop * 1.000001) + 1;
  if (tid < kGroups * 3) (&s_gmin[0][0])[tid] = 0x7fffffff;

  for (;;) {
    // ---- next source (device-side do_grid_master: one ticket per work group) ---------------------
    __syncthreads();
    if (tid < kGroups * 2) (&s_next[0][0])[tid] = 0;
    if (crank == 0 && tid == 0) s_work = (int)atomicAdd(P.ticket, 1u);
    int w;
    if (kCluster > 1) {
      cluster.sync();
      w = *cluster.map_shared_rank(&s_work, 0);
    } else {
      __syncthreads();
      w = s_work;
    }
    // the work list, then (single-CTA kernel only) the sources the per-warp kernel handed over after their first
    // subbox: their pass 1 is repeated without rates (`silent`), only to rebuild the planes of shell `subboxsize`
    const int nwork_all = P.nwork + ((kCluster == 1 && P.ovf_count) ? (int)*P.ovf_count : 0);
    if (w >= nwork_all) break;
    const bool handed_over = w >= P.nwork;
    const int ns = handed_over ? P.ovf[w - P.nwork] : P.work[w];  // 0-based source index
    const int silent = handed_over ? 1 : 0;                       // passes nbox <= silent emit nothing
    const int src0 = P.srcpos[3 * ns] - 1, src1 = P.srcpos[3 * ns + 1] - 1, src2 = P.srcpos[3 * ns + 2] - 1;
    S.normflux = P.normflux[ns];
    const double total_source_flux = S.normflux * P.S_star;  // evolve_source.F90:119

    int nbox = 0;
    double photon_loss_src = total_source_flux;  // :121
    int lr0 = 0, lr1 = 0, lr2 = 0, ll0 = 0, ll1 = 0, ll2 = 0;  // last_r-src, src-last_l per axis
    int r_done = -1;
    bool face_dead = false;
    // do while (evolve_source.F90:128-131); every thread of the work group evaluates it on identical values
    while (photon_loss_src > P.loss_fraction * total_source_flux && lr2 < P.lim[2][1] && ll2 < P.lim[2][0]) {
      nbox += 1;
      const int reach = P.subboxsize * nbox;  // :135-136
      S.reach = reach;
      S.emit = S.normflux > 0.0 && nbox > silent;
      lr0 = min(reach, P.lim[0][1]); ll0 = min(reach, P.lim[0][0]);
      lr1 = min(reach, P.lim[1][1]); ll1 = min(reach, P.lim[1][0]);
      lr2 = min(reach, P.lim[2][1]); ll2 = min(reach, P.lim[2][0]);
      const int rmax = max(max(max(lr0, ll0), max(lr1, ll1)), max(lr2, ll2));
      if (tid < kNf) {
        const int f = (int)crank * kNf + tid;
        s_face[tid] = P.use_twins ? make_face<true>(P, f, src0, src1, src2, lr0, ll0, lr1, ll1, lr2, ll2)
                                  : make_face<false>(P, f, src0, src1, src2, lr0, ll0, lr1, ll1, lr2, ll2);
      }
      double loss = 0.0;
      if (r_done < 0) {
        // shell 0 = the source cell (evolve_point.F90:151-160): coldensh_in=0, path=dr/2, vol_ph=cell volume.
        // Every quadrant stores it as its plane 0; the first thread of the +z face owns it.
        if (tid < 4 * kNf) {
          const unsigned cell = ((unsigned)src2 * (unsigned)P.n[1] + (unsigned)src1) * (unsigned)P.n[0] + (unsigned)src0;
          const double tau_cell = P.tau_cell[cell];
          const double tau_out = 0.5 * tau_cell;
          s_planes[(size_t)(2 * (tid >> 2)) * (cap + kPadFront) + kPadFront + (tid & 3)] = tau_out;   // plane 0 of face tid/4, quadrant tid%4
          if (crank == 0 && tid == 0) {
            if (kDebug) P.coldens_dbg[cell] = tau_out * P.inv_sigma;
            if (S.emit) {
              double phi_all, phi_out, heat_all = 0.0;
              photo_rates<kHeat>(0.0, tau_out, S.normflux, s_thick, s_logtab, s_heat, P, L, phi_all, phi_out, heat_all);
              // rate = phi_all/(vol_cell*nHI), nHI = tau_cell/(sigma*dr0)
              const double photo_cell = phi_all * fast_rcp(P.vol_cell * tau_cell * P.inv_sigma_dr0);
              if (photo_cell != 0.0) atomicAdd(&P.phih[cell], photo_cell);
              if (kHeat) {
                const double heat_cell = heat_all * fast_rcp(P.vol_cell);
                if (heat_cell != 0.0) atomicAdd(&P.phiheat[cell], heat_cell);
              }
            }
          }
        }
        r_done = 0;
      }
      __syncthreads();   // s_face and plane 0 visible
      // Faces never exchange data (a cell reads the previous plane of its own quadrant only), so the kWpf warps of
      // a face walk its shells on their own and the work group meets again only for the loss of the pass.
      for (int r = r_done + 1; r <= rmax; ++r) {
        int min_hi = 0x7fffffff;   // high word of the smallest optical depth this thread writes to plane r
        if (gtid == 0) s_next[g][(r + 1) & 1] = 0;   // the counter of shell r+1 (nobody uses it before the barrier below)
        if (!face_dead) {
          const int P1 = r + 1;
          const int nseg = (kCluster == 1) ? P.nseg_cta[r] : P.nseg_cl[r];   // host-tuned split of the columns along b
          // a plane of side s fits a shared buffer together with the slack the zero-weight reads may touch
          const bool cur_sm = 4 * P1 * P1 + 4 * P1 + 8 <= cap;
          const bool prev_sm = 4 * r * r + 4 * r + 8 <= cap;
          const Face* faces = s_face + fl;
          if (cur_sm) {
            double* cur = (r & 1) ? sbuf1 : sbuf0;
            const double* prev = (r & 1) ? sbuf0 : sbuf1;
            if (r == 1) trace_shell<kFpg, kTg, false, false, true, kLls, kDebug, kHeat>(P, S, faces, gtid, s_thick, s_logtab, s_heat, L, r, prev, cur, sstride, nseg, &s_next[g][r & 1], loss, min_hi);
            else if (r < S.rsafe) trace_shell<kFpg, kTg, false, false, false, kLls, kDebug, kHeat>(P, S, faces, gtid, s_thick, s_logtab, s_heat, L, r, prev, cur, sstride, nseg, &s_next[g][r & 1], loss, min_hi);
            else trace_shell<kFpg, kTg, false, true, false, kLls, kDebug, kHeat>(P, S, faces, gtid, s_thick, s_logtab, s_heat, L, r, prev, cur, sstride, nseg, &s_next[g][r & 1], loss, min_hi);
          } else {
            double* cur = (r & 1) ? gbuf1 : gbuf0;
            double* gprev = (r & 1) ? gbuf0 : gbuf1;
            if (prev_sm) {
              // first shell that does not fit: the group moves its planes r-1 to the global scratch
              const double* sprev = (r & 1) ? sbuf0 : sbuf1;
              for (int f = 0; f < kFpg; ++f)
                for (int i = gtid; i < 4 * r * r; i += kTg) gprev[f * gstride + i] = sprev[f * sstride + i];
              group_sync<kTg>(g);
            }
            if (r == 1) trace_shell<kFpg, kTg, true, false, true, kLls, kDebug, kHeat>(P, S, faces, gtid, s_thick, s_logtab, s_heat, L, r, gprev, cur, gstride, nseg, &s_next[g][r & 1], loss, min_hi);
            else if (r < S.rsafe) trace_shell<kFpg, kTg, true, false, false, kLls, kDebug, kHeat>(P, S, faces, gtid, s_thick, s_logtab, s_heat, L, r, gprev, cur, gstride, nseg, &s_next[g][r & 1], loss, min_hi);
            else trace_shell<kFpg, kTg, true, true, false, kLls, kDebug, kHeat>(P, S, faces, gtid, s_thick, s_logtab, s_heat, L, r, gprev, cur, gstride, nseg, &s_next[g][r & 1], loss, min_hi);
          }
        }
        // plane r complete before plane r+1 reads it.  Dead-face rule: once every optical depth of the planes of a
        // group exceeds the 2e19 column (evolve_point.F90:201) by a margin, every later cell of these faces
        // interpolates to more than the limit as well (a weighted mean of such values, plus non-negative terms):
        // no rate, no boundary loss, so the group skips the arithmetic of its remaining shells (the cells still
        // count as updates).
        const int wmin = __reduce_min_sync(0xffffffffu, min_hi);
        if (kTg == 32) {
          __syncwarp();
          if (!kDebug && !face_dead) face_dead = wmin > dead_hi;
        } else {
          const int slot3 = r % 3;
          if (lane == 0 && !face_dead) atomicMin(&s_gmin[g][slot3], wmin);
          group_sync<kTg>(g);
          if (!kDebug && !face_dead) face_dead = s_gmin[g][slot3] > dead_hi;
          if (gtid == 0) s_gmin[g][(r + 2) % 3] = 0x7fffffff;   // the slot of shell r+2: nobody touches it before the barrier of r+1
        }
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
        if (nbox > silent) photon_loss_src = t;            // (a silent pass keeps the loop going: its loss was > the cut-off)
      } else {
        if (tid == 0) {
          double t = 0.0;
          for (int i = 0; i < kT / 32; ++i) t += s_red[i];
          cluster.map_shared_rank(&s_slot[pass_parity][0], 0)[crank] = t;   // DSMEM store into rank 0
        }
        cluster.sync();
        const double* slots = cluster.map_shared_rank(&s_slot[pass_parity][0], 0);
        double t = 0.0;
        for (int i = 0; i < kCluster; ++i) t += slots[i];   // fixed order: identical in all CTAs
        photon_loss_src = t;
        pass_parity ^= 1;
      }
      __syncthreads();   // s_red / s_face are rewritten by the next pass
    }
    if (crank == 0 && tid == 0) {
      P.nbox_out[ns] = nbox;              // sum_nbox=sum_nbox+nbox, :219
      P.loss_out[ns] = photon_loss_src;   // photon_loss(1)=photon_loss(1)+photon_loss_src, :216
    }
    if (kCluster > 1) cluster.sync();     // every peer has read s_work before rank 0 fetches the next ticket
  }
  if (kCluster > 1) cluster.sync();  // nobody leaves while a peer may still read rank 0's shared memory
}

// The first subbox of many short traces: ONE WARP per source, kWarpMax warps per CTA, one CTA per SM.
// Early in reionization every trace ends after one or two subboxes (a few hundred to a few thousand cells); a whole CTA
// per source then spends its time in barriers and in the latency chain ticket -> source -> first loads, with two
// sources in flight per SM.  Here a warp walks shells 0..subboxsize of its source alone (planes in its own slice of
// shared memory, __syncwarp between shells), so an SM has up to 10-12 independent traces in flight.  A source whose
// loss after this first subbox still exceeds the cut-off (evolve_source.F90:128-131) is handed over: the single-CTA
// kernel, launched next, repeats its pass 1 without rates and carries on from pass 2.  The host routes a source here
// only when its previous trace ended after one subbox, and only if subboxsize < every half-box limit (no clipping).
constexpr int kWarpMax = 12;

template <int kLls, bool kHeat>
__global__ void __launch_bounds__(32 * kWarpMax, 1) raytrace_warp_kernel(RtParams P) {
  extern __shared__ double2 smem2[];
  double2* s_thick = smem2;
  double2* s_logtab = smem2 + kTableLen;
  double2* s_heat = smem2 + kTableLen + 128;
  double* s_planes = reinterpret_cast<double*>(smem2 + kTableLen + 128 + (kHeat ? kTableLen : 0));
  __shared__ Face s_face[kWarpMax][kFaces];
  __shared__ int s_next[kWarpMax];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nthreads = blockDim.x;
  const int cap0 = P.warp_plane_doubles[0], cap1 = P.warp_plane_doubles[1];   // planes of the even / odd shells
  for (int i = tid; i < kTableLen; i += nthreads) s_thick[i] = P.thick2[i];
  for (int i = tid; i < 128; i += nthreads) s_logtab[i] = P.logtab[i];
  if (kHeat)
    for (int i = tid; i < kTableLen; i += nthreads) s_heat[i] = P.heat2[i];
  const int sstride = cap0 + cap1 + 2 * kPadFront;   // from a face's two buffers to the next face's
  const int per_warp = kFaces * sstride;
  for (int i = tid; i < (nthreads >> 5) * per_warp; i += nthreads) s_planes[i] = 0.0;
  __syncthreads();
  double* wbase = s_planes + (size_t)wid * per_warp;
  double* sbuf0 = wbase + kPadFront;
  double* sbuf1 = sbuf0 + cap0 + kPadFront;
  LogC L;
  L.c0 = P.logc[0]; L.c1 = P.logc[1]; L.c2 = P.logc[2]; L.c3 = P.logc[3]; L.c4 = P.logc[4]; L.B = P.logB;
  SrcCtx S;
  S.rsafe = min(min(min(P.lim[0][0], P.lim[0][1]), min(P.lim[1][0], P.lim[1][1])), min(P.lim[2][0], P.lim[2][1]));
  const int reach = P.subboxsize;   // nbox = 1; the host guarantees reach < rsafe
  S.reach = reach;
  const int dead_hi = __double2hiint(P.tau_stop * 1.000001) + 1;
  for (;;) {
    int w = 0;
    if (lane == 0) w = (int)atomicAdd(P.ticket, 1u);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= P.nwork) break;
    const int ns = P.work[w];
    const int src0 = P.srcpos[3 * ns] - 1, src1 = P.srcpos[3 * ns + 1] - 1, src2 = P.srcpos[3 * ns + 2] - 1;
    S.normflux = P.normflux[ns];
    S.emit = S.normflux > 0.0;
    const double total_source_flux = S.normflux * P.S_star;   // evolve_source.F90:119
    if (!(total_source_flux > P.loss_fraction * total_source_flux)) {   // the do while of :128 is never entered
      if (lane == 0) {
        P.nbox_out[ns] = 0;
        P.loss_out[ns] = total_source_flux;
      }
      continue;
    }
    __syncwarp();   // the lanes are done with the previous source's faces and planes
    if (lane < kFaces) s_face[wid][lane] = make_face<false>(P, lane, src0, src1, src2, reach, reach, reach, reach, reach, reach);
    if (lane < 4 * kFaces) {
      // shell 0 = the source cell (evolve_point.F90:151-160), plane 0 of every quadrant
      const unsigned cell = ((unsigned)src2 * (unsigned)P.n[1] + (unsigned)src1) * (unsigned)P.n[0] + (unsigned)src0;
      const double tau_cell = P.tau_cell[cell];
      const double tau_out = 0.5 * tau_cell;
      sbuf0[(lane >> 2) * sstride + (lane & 3)] = tau_out;
      if (lane == 0) {
        double phi_all, phi_out, heat_all = 0.0;
        photo_rates<kHeat>(0.0, tau_out, S.normflux, s_thick, s_logtab, s_heat, P, L, phi_all, phi_out, heat_all);
        const double photo_cell = phi_all * fast_rcp(P.vol_cell * tau_cell * P.inv_sigma_dr0);
        if (photo_cell != 0.0) atomicAdd(&P.phih[cell], photo_cell);
        if (kHeat) {
          const double heat_cell = heat_all * fast_rcp(P.vol_cell);
          if (heat_cell != 0.0) atomicAdd(&P.phiheat[cell], heat_cell);
        }
      }
    }
    double loss = 0.0;
    for (int r = 1; r <= reach; ++r) {
      int min_hi = 0x7fffffff;   // high word of the smallest optical depth written to plane r
      if (lane == 0) s_next[wid] = 0;
      __syncwarp();   // plane r-1, the faces and the work counter are visible to the warp
      double* cur = (r & 1) ? sbuf1 : sbuf0;
      const double* prev = (r & 1) ? sbuf0 : sbuf1;
      const int nseg = P.nseg_w[r];
      if (r == 1) trace_shell<kFaces, 32, false, false, true, kLls, false, kHeat>(P, S, s_face[wid], lane, s_thick, s_logtab, s_heat, L, r, prev, cur, sstride, nseg, &s_next[wid], loss, min_hi);
      else trace_shell<kFaces, 32, false, false, false, kLls, false, kHeat>(P, S, s_face[wid], lane, s_thick, s_logtab, s_heat, L, r, prev, cur, sstride, nseg, &s_next[wid], loss, min_hi);
      // dead-shell rule (see raytrace_kernel): once every optical depth of plane r exceeds the 2e19 column by a
      // margin, no later cell of this source has a rate or a boundary loss; the cells still count as updates
      if (__reduce_min_sync(0xffffffffu, min_hi) > dead_hi) break;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0) {
      // do while (photon_loss_src > loss_fraction*total_source_flux ...), evolve_source.F90:128-131 (the subbox is
      // inside the half box on every side, so only the loss decides)
      if (loss > P.loss_fraction * total_source_flux) {
        P.ovf[atomicAdd(P.ovf_count, 1u)] = ns;   // hand over: the single-CTA kernel carries on from pass 2
      } else {
        P.nbox_out[ns] = 1;
        P.loss_out[ns] = loss;
      }
    }
  }
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
  // one work-group slot: two plane buffers of six faces
  return (size_t)2 * kFaces * raytrace_face_doubles(plane_stride);
}

// plane_doubles = capacity of one plane buffer of one face; a CTA holds two per face
static size_t rt_smem_bytes(int plane_doubles, int nfaces, bool heat) {
  return (size_t)(kTableLen + 128 + (heat ? kTableLen : 0)) * sizeof(double2) +
         (size_t)2 * nfaces * (plane_doubles + kPadFront) * sizeof(double);
}

typedef void (*RtKernel)(RtParams);

// lls: 0/1 = scalar (0: no LLS, a zero column), 2 = LLS_grid, 3 = R_max; the diagnostic (debug) kernels exist for
// the isothermal path only
template <int kCluster>
static RtKernel pick_kernel(int lls, bool debug, bool heat) {
  if (debug) {
    switch (lls) {
      case 2: return raytrace_kernel<kCluster, 2, true, false>;
      case 3: return raytrace_kernel<kCluster, 3, true, false>;
      default: return raytrace_kernel<kCluster, 1, true, false>;
    }
  }
  if (heat) {
    switch (lls) {
      case 2: return raytrace_kernel<kCluster, 2, false, true>;
      case 3: return raytrace_kernel<kCluster, 3, false, true>;
      default: return raytrace_kernel<kCluster, 1, false, true>;
    }
  }
  switch (lls) {
    case 2: return raytrace_kernel<kCluster, 2, false, false>;
    case 3: return raytrace_kernel<kCluster, 3, false, false>;
    default: return raytrace_kernel<kCluster, 1, false, false>;
  }
}

static void cluster_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int nclusters, size_t smem,
                           cudaStream_t stream) {
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3(kClusterSize * nclusters, 1, 1);
  cfg->blockDim = dim3(kT, 1, 1);
  cfg->dynamicSmemBytes = smem;
  cfg->stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kClusterSize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
}

static RtKernel pick_warp_kernel(int lls, bool heat) {
  if (heat) {
    switch (lls) {
      case 2: return raytrace_warp_kernel<2, true>;
      case 3: return raytrace_warp_kernel<3, true>;
      default: return raytrace_warp_kernel<1, true>;
    }
  }
  switch (lls) {
    case 2: return raytrace_warp_kernel<2, false>;
    case 3: return raytrace_warp_kernel<3, false>;
    default: return raytrace_warp_kernel<1, false>;
  }
}

int raytrace_configure(int max_radius, bool heat_tables, int subboxsize, int min_lim, RtLaunchInfo* info) {
  int dev = 0, max_optin = 0, sm_total = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&sm_total, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t fixed = (size_t)(kTableLen + 128 + (heat_tables ? kTableLen : 0)) * sizeof(double2);
  // kCtaPerSm CTAs per SM share the opt-in shared memory; what the tables leave goes to the plane buffers
  int per_cta = std::min(std::min(max_optin, sm_total / kCtaPerSm - 2048), (heat_tables ? 88 : 56) * 1024);   // the rest of the SM's 256 KB is L1 for the global planes
  if (const char* env = getenv("C2B_RT_SMEM_KB")) per_cta = std::min(std::min(max_optin, sm_total / kCtaPerSm - 2048), atoi(env) * 1024);
  const int avail = (int)(((size_t)per_cta - fixed - 1024) / sizeof(double));   // doubles for all plane buffers of a CTA
  const int full = 4 * (max_radius + 1) * (max_radius + 1) + 4 * (max_radius + 1) + 8;   // one face at the largest radius
  const int cap_cta = std::max(64, std::min(avail / (2 * kFaces) - kPadFront, full)) & ~3;
  const int cap_cl = std::max(64, std::min(avail / 2 - kPadFront, full)) & ~3;
  for (int var = 0; var < 3; ++var)   // 0: plain, 1: diagnostic, 2: with heating rates
    for (int lls = 1; lls < 4; ++lls) {
      const bool dbg = var == 1, heat = var == 2;
      cudaError_t e = cudaFuncSetAttribute(pick_kernel<1>(lls, dbg, heat), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)rt_smem_bytes(cap_cta, kFaces, heat));
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(pick_kernel<kClusterSize>(lls, dbg, heat), cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)rt_smem_bytes(cap_cl, 1, heat));
      if (e != cudaSuccess) return (int)e;
    }
  info->smem_plane_doubles = cap_cta;
  info->smem_plane_doubles_cl = cap_cl;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pick_kernel<1>(1, false, heat_tables), kT, rt_smem_bytes(cap_cta, kFaces, heat_tables));
  info->grid_cta = sms * std::max(1, per_sm);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  cluster_config(&cfg, attr, sms, rt_smem_bytes(cap_cl, 1, heat_tables), nullptr);
  int nclusters = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, pick_kernel<kClusterSize>(1, false, heat_tables), &cfg);
  if (e != cudaSuccess) return (int)e;
  info->clusters = std::max(1, nclusters);
  info->cluster_size = kClusterSize;
  // scratch slots: one per resident CTA of the single-CTA kernel plus one per resident cluster
  info->grid_max = info->grid_cta + info->clusters;
  // per-warp kernel: shells 0..subboxsize of one source per warp, planes of every warp in shared memory
  info->warp_warps = 0;
  info->warp_plane_doubles[0] = info->warp_plane_doubles[1] = 0;
  info->grid_warp = sms;
  if (subboxsize >= 1 && subboxsize < min_lim) {
    // shells alternate between two buffers per face: even r in the first, odd r in the second
    auto plane_cap = [](int r) { const int P1 = r + 1; return (4 * P1 * P1 + 4 * P1 + 8 + 3) & ~3; };
    const int r_even = subboxsize & ~1, r_odd = (subboxsize - 1) | 1;
    const int cap_e = plane_cap(r_even), cap_o = plane_cap(r_odd);
    const size_t per_warp = (size_t)kFaces * (cap_e + cap_o + 2 * kPadFront) * sizeof(double);
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, pick_warp_kernel(1, heat_tables)) != cudaSuccess) return (int)cudaGetLastError();
    const long room = (long)max_optin - (long)fa.sharedSizeBytes - (long)fixed - 256;
    int nw = room > 0 ? (int)std::min<long>(kWarpMax, room / (long)per_warp) : 0;
    if (const char* env = getenv("C2B_WARP_WARPS")) nw = std::min(nw, atoi(env));
    if (nw >= 4) {
      for (int lls = 1; lls < 4; ++lls)
        for (int heat = 0; heat < 2; ++heat) {
          const size_t fx = (size_t)(kTableLen + 128 + (heat ? kTableLen : 0)) * sizeof(double2);
          cudaError_t e2 = cudaFuncSetAttribute(pick_warp_kernel(lls, heat != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)std::min<size_t>(max_optin - fa.sharedSizeBytes, fx + kWarpMax * per_warp));
          if (e2 != cudaSuccess) return (int)e2;
        }
      info->warp_warps = nw;
      info->warp_plane_doubles[0] = cap_e;
      info->warp_plane_doubles[1] = cap_o;
    }
  }
  return 0;
}

// Number of b-segments per column for shell r: minimises rounds x (segment length + set-up) for the `threads`
// threads that walk a group of `nf` faces, i.e. nf*(r+1) columns of four cells per row.
static void build_nseg_table(int max_radius, int nf, int threads, std::vector<int>& tab) {
  tab.assign((size_t)max_radius + 2, 1);
  if (const char* env = getenv("C2B_RT_SEGLEN")) {   // development knob: fixed target segment length
    const int target = std::max(1, atoi(env));
    for (int r = 1; r <= max_radius; ++r) tab[r] = std::min(64, (r + 1 + target - 1) / target);
    return;
  }
  for (int r = 1; r <= max_radius; ++r) {
    const int P1 = r + 1, ncol = nf * P1;
    long best = -1;
    int best_n = 1;
    for (int n = 1; n <= P1 && n <= 64; ++n) {
      const long rounds = ((long)ncol * n + threads - 1) / threads;
      const long cost = rounds * (2 * ((P1 + n - 1) / n) + 3);   // a row costs ~2 set-ups' worth... set-up ~1.5 rows
      if (best < 0 || cost < best) { best = cost; best_n = n; }
    }
    tab[r] = best_n;
  }
}

void raytrace_nseg_tables(int max_radius, std::vector<int>& cta, std::vector<int>& cl, std::vector<int>& warp) {
  build_nseg_table(max_radius, kFpgCta, kTgCta, cta);   // single-CTA kernel: groups of kFpgCta faces
  build_nseg_table(max_radius, 1, kT, cl);              // cluster kernel: the whole CTA walks one face
  build_nseg_table(max_radius, kFaces, 32, warp);       // per-warp kernel: one warp walks the six faces
}

static int lls_mode(const RtParams& p) { return p.use_lls ? p.type_lls : 0; }

void launch_raytrace(const RtParams& p, int grid, cudaStream_t stream) {
  const bool heat = p.phiheat != nullptr && p.coldens_dbg == nullptr;   // the diagnostic kernels carry no heating rates
  pick_kernel<1>(lls_mode(p), p.coldens_dbg != nullptr, heat)<<<grid, kT, rt_smem_bytes(p.smem_plane_doubles, kFaces, heat), stream>>>(p);
}

void launch_raytrace_warp(const RtParams& p, int grid, int warps, cudaStream_t stream) {
  const bool heat = p.phiheat != nullptr;
  const size_t smem = (size_t)(kTableLen + 128 + (heat ? kTableLen : 0)) * sizeof(double2) +
                      (size_t)warps * kFaces * (p.warp_plane_doubles[0] + p.warp_plane_doubles[1] + 2 * kPadFront) * sizeof(double);
  pick_warp_kernel(lls_mode(p), heat)<<<grid, 32 * warps, smem, stream>>>(p);
}

int launch_raytrace_cluster(const RtParams& p, int nclusters, cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  const bool heat = p.phiheat != nullptr && p.coldens_dbg == nullptr;
  cluster_config(&cfg, attr, nclusters, rt_smem_bytes(p.smem_plane_doubles_cl, 1, heat), stream);
  return (int)cudaLaunchKernelEx(&cfg, pick_kernel<kClusterSize>(lls_mode(p), p.coldens_dbg != nullptr, heat), p);
}

void launch_taucell(const float* ndens, const double* xh_av, double* tau_cell, size_t n, double sigma_dr0,
                    double eps, cudaStream_t stream) {
  taucell_kernel<<<chemistry_blocks(), 256, 0, stream>>>(ndens, xh_av, tau_cell, n, sigma_dr0, eps);
}

void launch_pair_table(const double* tab, double2* out, cudaStream_t stream) {
  pair_table_kernel<<<(kTableLen + 127) / 128, 128, 0, stream>>>(tab, out);
}

}  // namespace c2b
