// Spectral differential operators (banded mat-vec) and the semi-implicit solves (row-pivoted banded LU),
// one (m,k) column per CTA / warp.  Reference: /root/reference/src/submodules/mlegs_scalar_ops.f90:418-1000
// with the coefficient formulas of mlegs_spectfm_diff.f90:6-152 and LAPACK's zgbtf2/zgbtrs/ztbsv
// operation order (external/lapack/SRC/zgbtf2.f:219-262, zgbtrs.f:205-232).
//
// The reference rebuilds a dense nn x nn matrix with nn^2 exp() calls for every (m,k); here the
// lognorm-scaled band coefficients are tabulated once per m on the host (same formulas, same libm) and
// the -k^2 shift is applied on the fly.  Compiled with -fmad=false: products and sums round separately,
// in the reference's order, so matrices are bit-identical to a non-fused CPU evaluation.
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "kernels.h"

namespace mlegs {

// ---------------------------------------------------------------------------------------------
// host tables: leg_xxdx (sdiff:6-43) and leg_del2h (sdiff:45-96) for the untruncated band
// ---------------------------------------------------------------------------------------------
int build_operator_tables() {
  Context &c = ctx();
  const int ne = c.ne, nm = c.p.npchop;
  const double ell2 = std::pow(c.p.ell, 2.0);
  c.h_del2h.assign((size_t)nm * 5 * ne, 0.0);
  c.h_xxdx.assign((size_t)nm * 3 * ne, 0.0);
  const double s = 0.0;
  for (int am = 0; am < nm; ++am) {
    const double *ln = c.h_lognorm.data() + (size_t)am * ne;
    double *t5 = c.h_del2h.data() + (size_t)am * 5 * ne;
    double *t3 = c.h_xxdx.data() + (size_t)am * 3 * ne;
    for (int i = 0; i < ne; ++i) {
      const double n = (double)(am + i);
      const double nam = (double)(i);   // n - am
      double v[5];
      v[0] = -(n - am - 1.0) * nam * (n - 2.0 + s) * (n - 1.0 + s) / (2.0 * n - 3.0) / (2.0 * n - 1.0);
      v[1] = 2.0 * n * nam * (n - 1.0 + s) / (2.0 * n - 1.0);
      v[2] = (-2.0 * n * (n + 1.0) * (3.0 * n * n + 3.0 * n - (double)(am * am) - 2.0) +
              2.0 * s * (s - 2.0) * (n * n + n + (double)(am * am) - 1.0)) /
             (2.0 * n - 1.0) / (2.0 * n + 3.0);
      v[3] = 2.0 * (n + 1.0) * (n + am + 1.0) * (n + 2.0 - s) / (2.0 * n + 3.0);
      v[4] = -(n + am + 1.0) * (n + am + 2.0) * (n + 3.0 - s) * (n + 2.0 - s) / (2.0 * n + 3.0) / (2.0 * n + 5.0);
      for (int b = 0; b < 5; ++b) {
        int j = i + b - 2;
        if (j < 0 || j >= ne) continue;
        double g = v[b] / ell2;
        t5[(size_t)b * ne + i] = g * std::exp(ln[j] - ln[i]);
      }
      double x[3];
      x[0] = -(n - 1.0) * nam / (2.0 * n - 1.0);
      x[1] = 0.0;
      x[2] = (n + 2.0) * (n + am + 1.0) / (2.0 * n + 3.0);
      for (int b = 0; b < 3; ++b) {
        int j = i + b - 1;
        if (j < 0 || j >= ne || b == 1) continue;
        t3[(size_t)b * ne + i] = x[b] * std::exp(ln[j] - ln[i]);
      }
    }
  }
  CUDA_TRY(cudaMalloc((void **)&c.d_del2h, c.h_del2h.size() * sizeof(double)));
  CUDA_TRY(cudaMemcpy(c.d_del2h, c.h_del2h.data(), c.h_del2h.size() * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc((void **)&c.d_xxdx, c.h_xxdx.size() * sizeof(double)));
  CUDA_TRY(cudaMemcpy(c.d_xxdx, c.h_xxdx.data(), c.h_xxdx.size() * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc((void **)&c.d_flag, 4 * sizeof(int)));
  CUDA_TRY(cudaMemset(c.d_flag, 0, 4 * sizeof(int)));

  // vec2tp projection tables (ops:1405-1410): v = pf/n/(n+1)/fff, d = (pf .mul. xxdx)/n/(n+1)/fff
  const int nrh = c.nrh, nr = c.p.nr;
  std::vector<double> hv((size_t)nrh * ne * nm, 0.0), hd((size_t)nrh * ne * nm, 0.0);
  std::vector<double> hpf((size_t)nrh * ne * nm);
  CUDA_TRY(cudaMemcpy(hpf.data(), c.d_pf, hpf.size() * sizeof(double), cudaMemcpyDeviceToHost));
  std::vector<double> fff(nrh);
  for (int i = 0; i < nrh; ++i) fff[i] = (1.0 - std::pow(c.h_x[i], 2.0)) / c.h_w[i];
  (void)nr;
  unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nthreads; ++t) {
    pool.emplace_back([&, t]() {
      for (int m = (int)t; m < nm; m += (int)nthreads) {
        const double *pf = hpf.data() + (size_t)m * nrh * ne;
        const double *t3 = c.h_xxdx.data() + (size_t)m * 3 * ne;
        double *v = hv.data() + (size_t)m * nrh * ne;
        double *d = hd.data() + (size_t)m * nrh * ne;
        for (int col = 0; col + 1 < ne; ++col) {
          int nq = std::max(1, m + col);
          for (int i = 0; i < nrh; ++i) {
            // pfd(:,col) = sum_k pf(:,k) X(k,col), k ascending (bops:135-160); X(k,col) = t3[(col-k+1)][k]
            double acc = 0.0;
            if (col - 1 >= 0) acc = acc + pf[(size_t)(col - 1) * nrh + i] * t3[(size_t)2 * ne + (col - 1)];
            acc = acc + pf[(size_t)col * nrh + i] * 0.0;
            acc = acc + pf[(size_t)(col + 1) * nrh + i] * t3[(size_t)0 * ne + (col + 1)];
            v[(size_t)col * nrh + i] = pf[(size_t)col * nrh + i] / nq / (nq + 1) / fff[i];
            d[(size_t)col * nrh + i] = acc / nq / (nq + 1) / fff[i];
          }
        }
      }
    });
  }
  for (auto &th : pool) th.join();
  CUDA_TRY(cudaMalloc((void **)&c.d_vtab, hv.size() * sizeof(double)));
  CUDA_TRY(cudaMemcpy(c.d_vtab, hv.data(), hv.size() * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc((void **)&c.d_dtab, hd.size() * sizeof(double)));
  CUDA_TRY(cudaMemcpy(c.d_dtab, hd.data(), hd.size() * sizeof(double), cudaMemcpyHostToDevice));
  return MLEGS_OK;
}

// ---------------------------------------------------------------------------------------------
// banded mat-vec: xxdx / del2h / del2 / helmp
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int nn_of(int mglob, int nrc, int npc) {
  if (mglob >= npc) return 0;
  int v = min(nrc, nrc - mglob);
  return v > 0 ? v : 0;
}

// how many times the reference's two axial loops visit plane k (0-based); see k_ranges in the oracle
__device__ __forceinline__ int k_visits(int k, int nzl, int nzc, int nzcu) {
  int v = 0;
  if (k < min(nzl, nzc)) v++;
  int lo = max(nzcu, 1) - 1;
  if (k >= lo && k < nzl) v++;
  return v;
}

// One CTA per (column m, block of KB axial planes).  A thread owns one radial row i for all KB planes, so its band
// coefficients are loaded once and reused KB times; planes are staged in shared memory ([plane][row], row fastest:
// conflict-free, coalesced), which also makes the in-place update safe.  The -k^2 shift of the diagonal is applied
// per plane.  `reps` keeps the reference's habit of visiting the plane k = nz/2 twice when nzchop = nz/2 + 1.
#define BAND_KB_MAX 8
template <int NB>
__global__ void __launch_bounds__(1024) band_op_kernel(BandOpArgs a, int KB) {
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx *bufs = reinterpret_cast<cplx *>(smraw);
  const size_t bsz = (size_t)KB * a.nrl;
  cplx *cur = bufs, *nxt = bufs + bsz, *s0 = bufs + 2 * bsz, *s2 = bufs + 3 * bsz;   // s0/s2 only with combine
  const int j = blockIdx.x, k0 = blockIdx.y * KB;
  const int mglob = a.m0 + j * a.ms;
  const int nn = nn_of(mglob, a.nrc, a.npc);
  constexpr int half = NB / 2;
  const double *tab = a.tab + (size_t)mglob * NB * a.ne;
  __shared__ int s_reps[BAND_KB_MAX];
  __shared__ double s_ak2[BAND_KB_MAX];
  __shared__ int s_any, s_max;
  if (threadIdx.x < KB) {
    const int k = k0 + threadIdx.x;
    int reps = (k < a.nzl && nn >= 1) ? k_visits(k, a.nzl, a.nzc, a.nzcu) : 0;
    s_reps[threadIdx.x] = reps;
    double akv = (a.ak && k < a.nzl) ? a.ak[k] : 0.0;
    s_ak2[threadIdx.x] = akv * akv;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int mx = 0;
    for (int kk = 0; kk < KB; ++kk) mx = max(mx, s_reps[kk]);
    s_max = mx;
    s_any = (mx > 0) || a.combine || (mglob == 0 && k0 == 0 && a.nlnc > 0) || a.src != nullptr || a.out_neg;
  }
  __syncthreads();
  if (!s_any) return;
  const int maxreps = s_max;
  const bool lnblk = (mglob == 0 && k0 == 0 && a.nlnc > 0);

  // ---- stage the planes ----
  for (int kk = 0; kk < KB; ++kk) {
    const int k = k0 + kk;
    if (k >= a.nzl) break;
    const cplx *col = (a.src ? a.src : a.e) + ((size_t)k * a.npl + j) * a.nrl;
    for (int i = threadIdx.x; i < a.nrl; i += blockDim.x) {
      cplx v = col[i];
      cur[kk * a.nrl + i] = v;
      if (a.combine) s0[kk * a.nrl + i] = v;
    }
  }
  __syncthreads();

  const bool every_line = a.src != nullptr || a.out_neg;   // out of place / negated: lines the operator leaves alone are written too
  const bool direct = (a.napply == 1 && !a.combine && maxreps <= 1 && !lnblk && !every_line);
  for (int ap = 0; ap < a.napply; ++ap) {
    for (int r = 0; r < maxreps; ++r) {
      for (int i = threadIdx.x; i < a.nrl; i += blockDim.x) {
        double cf[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const int jj = i + b - half;
          cf[b] = (i < nn && jj >= 0 && jj < nn) ? __ldg(&tab[(size_t)b * a.ne + i]) : 0.0;
        }
        for (int kk = 0; kk < KB; ++kk) {
          const int k = k0 + kk;
          if (k >= a.nzl) break;
          const cplx *c0 = cur + kk * a.nrl;
          cplx o = c0[i];
          if (i < nn && r < s_reps[kk]) {
            double ar = 0.0, ai = 0.0;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const int jj = i + b - half;
              if (jj < 0 || jj >= nn) continue;
              double c = cf[b];
              if (b == half && a.ak) c = c - s_ak2[kk];
              cplx x = c0[jj];
              ar = ar + x.x * c;
              ai = ai + x.y * c;
            }
            o = make_double2(ar, ai);
            if (direct) a.e[((size_t)k * a.npl + j) * a.nrl + i] = o;
          }
          if (!direct) nxt[kk * a.nrl + i] = o;
        }
      }
      if (direct) return;
      __syncthreads();
      cplx *t = cur;
      cur = nxt;
      nxt = t;
    }
    if (ap == 0) {
      if (lnblk) {
        if (threadIdx.x < a.nlnc) {   // plane k = 0
          const double add = threadIdx.x == 0 ? a.lnc[0] : (threadIdx.x == 1 ? a.lnc[1] : a.lnc[2]);
          cur[threadIdx.x].x = cur[threadIdx.x].x + add;
        }
        __syncthreads();
      }
      if (a.combine) {
        for (int kk = 0; kk < KB; ++kk)
          for (int i = threadIdx.x; i < a.nrl; i += blockDim.x) s2[kk * a.nrl + i] = cur[kk * a.nrl + i];
        __syncthreads();
      }
    }
  }
  for (int kk = 0; kk < KB; ++kk) {
    const int k = k0 + kk;
    if (k >= a.nzl) break;
    if (!a.combine && !every_line && s_reps[kk] == 0 && !(lnblk && kk == 0)) continue;   // untouched plane
    cplx *col = a.e + ((size_t)k * a.npl + j) * a.nrl;
    for (int i = threadIdx.x; i < a.nrl; i += blockDim.x) {
      cplx o = cur[kk * a.nrl + i];
      if (a.combine) {
        cplx q = s2[kk * a.nrl + i], s = s0[kk * a.nrl + i];
        o = make_double2((o.x + a.beta * q.x) + a.alpha * s.x, (o.y + a.beta * q.y) + a.alpha * s.y);
      }
      if (a.out_neg) o = make_double2(-o.x, -o.y);
      col[i] = o;
    }
  }
}

int launch_band_op(const BandOpArgs &a, cudaStream_t st) {
  if (a.npl <= 0 || a.nzl <= 0) return MLEGS_OK;
  const int nbuf = a.combine ? 4 : 2;
  int KB = BAND_KB_MAX;
  static const char *cap_env = getenv("MLEGS_BAND_SMEM_KB");   // A/B: shared memory per CTA (planes per CTA vs CTAs per SM)
  const size_t cap = (size_t)(cap_env ? atoi(cap_env) : 36) * 1024;   // 256^3: xxdx 0.43 -> 0.31 ms, del2 0.35 -> 0.31 (72 -> 36 KB)
  while (KB > 1 && (size_t)nbuf * KB * a.nrl * sizeof(cplx) > cap) KB >>= 1;
  size_t smem = (size_t)nbuf * KB * a.nrl * sizeof(cplx);
  if (smem > 220 * 1024) return fail(MLEGS_E_ARG, "band operator: radial size too large for shared memory");
  static size_t attr_set3 = 0, attr_set5 = 0;
  if (a.nb == 3 && smem > 48 * 1024 && smem > attr_set3) {
    CUDA_TRY(cudaFuncSetAttribute(band_op_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set3 = smem;
  }
  if (a.nb == 5 && smem > 48 * 1024 && smem > attr_set5) {
    CUDA_TRY(cudaFuncSetAttribute(band_op_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set5 = smem;
  }
  dim3 grid(a.npl, (a.nzl + KB - 1) / KB);
  int threads = std::min(1024, (a.nrl + 31) / 32 * 32);
  // one pass over the retained lines (helmp's whole-array combination touches every line of the block)
  const double lines_b = (a.combine || a.src || a.out_neg) ? (double)a.nrl * a.npl * a.nzl
                                   : retained_elems(a.nrl, a.npl, a.nzl, 0, a.m0, a.nrc, a.npc, a.nzc, a.nzcu, a.ms);
  prof_begin(a.combine ? "helmp_band" : (a.nb == 3 ? "xxdx_band" : "del2_band"), st, 32.0 * lines_b);
  if (a.nb == 3)
    band_op_kernel<3><<<grid, threads, smem, st>>>(a, KB);
  else
    band_op_kernel<5><<<grid, threads, smem, st>>>(a, KB);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// ---------------------------------------------------------------------------------------------
// banded solves: ihelm / idel2 / ihelmp.  One warp per (m,k) system.
// workspace per system (doubles): D[5 nn] | Pa[17 nn] | AB[ldab nn] (aliases Pb) | rhs[2 nn]
// ---------------------------------------------------------------------------------------------
#define SOLVE_WARPS 4

__device__ __forceinline__ void warp_argmax_first(double &v, int &idx) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, off);
    int oi = __shfl_xor_sync(0xffffffffu, idx, off);
    if (ov > v || (ov == v && oi < idx)) {
      v = ov;
      idx = oi;
    }
  }
}

// P_out = D * P_in for nn x nn band matrices; P has half-width hw_in (bands -hw_in..hw_in, stored [(dc+HW)*nn + r]).
__device__ void band_product(const double *D, const double *Pin, int pin_off, double *Pout, int nn, int hw_in,
                             int HW, int lane) {
  const int hw_out = hw_in + 2;
  const int total = (2 * hw_out + 1) * nn;
  for (int idx = lane; idx < total; idx += 32) {
    int dc = idx / nn - hw_out;
    int r = idx - (dc + hw_out) * nn;
    int jc = r + dc;
    double acc = 0.0;
    if (jc >= 0 && jc < nn) {
      for (int da = -2; da <= 2; ++da) {
        int kk = r + da;
        int db = dc - da;
        if (kk < 0 || kk >= nn || db < -hw_in || db > hw_in) continue;
        acc = acc + D[(da + 2) * nn + r] * Pin[(db + pin_off) * nn + kk];
      }
    }
    Pout[(dc + HW) * nn + r] = acc;
  }
}

__global__ void __launch_bounds__(SOLVE_WARPS * 32) band_solve_kernel(SolveArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nsys = a.npl * a.nk;
  const int kl = a.kl, ku = a.ku, kv = kl + ku, ldab = 2 * kl + ku + 1;
  const int HW = 8;
  for (int sys = blockIdx.x * SOLVE_WARPS + warp; sys < nsys; sys += gridDim.x * SOLVE_WARPS) {
    const int j = sys % a.npl;
    const int k = a.k0 + sys / a.npl;
    const int mglob = a.m0 + j * a.ms;
    const int nn = nn_of(mglob, a.nrc, a.npc);
    if (nn < 1) continue;
    double *ws = a.ws_global ? a.ws_global + (size_t)(blockIdx.x * SOLVE_WARPS + warp) * a.ws_doubles
                             : reinterpret_cast<double *>(smraw) + (size_t)warp * a.ws_doubles;
    double *D = ws;
    double *Pa = D + 5 * a.nnmax;
    double *AB = Pa + (a.power > 2 ? 17 * a.nnmax : 0);
    double *Pb = AB;
    double *rhs = AB + (size_t)ldab * a.nnmax;
    int *piv = reinterpret_cast<int *>(rhs + 2 * a.nnmax);
    cplx *col = a.e + ((size_t)k * a.npl + j) * a.nrl;
    const double *tab = a.tab + (size_t)mglob * 5 * a.ne;
    const double akv = a.ak[k];
    const double ak2 = akv * akv;
    const bool special = (a.special00 && mglob == 0 && k == 0);

    // ---- D = leg_del2(m, ak(k), nn, nn) ----
    for (int idx = lane; idx < 5 * nn; idx += 32) {
      int b = idx / nn, r = idx - b * nn;
      int jc = r + b - 2;
      double cf = 0.0;
      if (jc >= 0 && jc < nn) {
        cf = tab[(size_t)b * a.ne + r];
        if (b == 2) cf = cf - ak2;
      }
      D[b * nn + r] = cf;
    }
    for (int idx = lane; idx < ldab * nn; idx += 32) AB[idx] = 0.0;
    // rhs (shifted down by one row for the prescribed-ln variant, ops:621-622)
    for (int i = lane; i < nn; i += 32) {
      cplx v;
      if (special && a.special00 == 2)
        v = (i == 0) ? make_double2(a.preln_rhs, 0.0) : col[i - 1];
      else
        v = col[i];
      rhs[2 * i] = v.x;
      rhs[2 * i + 1] = v.y;
    }
    __syncwarp();

    // ---- operator matrix into LAPACK band storage AB(kv + i - j, j) ----
    const double *H = D;
    int hwH = 2, HWs = 2;   // H stored with offset HWs
    if (a.power > 2) {
      // helmp_bnd = del2 .mul. helmp_bnd, power/2-1 times (ops:958-962).  Pb aliases AB, so the buffers
      // alternate such that the last product lands in Pa.
      int nprod = a.power / 2 - 1;
      const double *src = D;
      int hw = 2, src_off = 2;
      double *dst = (nprod % 2 == 1) ? Pa : Pb;
      for (int q = 0; q < nprod; ++q) {
        band_product(D, src, src_off, dst, nn, hw, HW, lane);
        __syncwarp();
        src = dst;
        src_off = HW;
        hw += 2;
        dst = (dst == Pa) ? Pb : Pa;
      }
      H = src;   // == Pa
      hwH = hw;
      HWs = HW;
      __syncwarp();
      // AB aliases Pb, which is dead now
      for (int idx = lane; idx < ldab * nn; idx += 32) AB[idx] = 0.0;
      __syncwarp();
    }
    {
      const int nbands = 2 * hwH + 1;
      for (int idx = lane; idx < nbands * nn; idx += 32) {
        int dc = idx / nn - hwH;
        int r = idx - (dc + hwH) * nn;
        int jc = r + dc;
        if (jc < 0 || jc >= nn) continue;
        if (dc > ku || -dc > kl) continue;
        double v = H[(dc + HWs) * nn + r];
        if (a.power > 2) {
          double dv = (dc >= -2 && dc <= 2) ? D[(dc + 2) * nn + r] : 0.0;
          v = v + a.beta * dv;           // fullmat(helmp) + beta*fullmat(del2), ops:963
        }
        if (dc == 0 && a.add_alpha) v = v + a.alpha;
        AB[(kv - dc) + (size_t)jc * ldab] = v;
      }
      __syncwarp();
    }
    if (special && lane == 0) {
      if (a.special00 == 1) {
        // idel2_proln, ops:705-711: first column picks up the log-term's Laplacian
        AB[(kv + 0) + 0 * ldab] = AB[(kv + 0)] + a.sp0;
        AB[(kv + 1) + 0 * ldab] = AB[(kv + 1)] - a.sp1;
        AB[(kv + 2) + 0 * ldab] = AB[(kv + 2)] + a.sp2;
      }
    }
    if (special && a.special00 == 2) {
      // idel2_preln, ops:609-618: rows shifted down by one, first row = e_1^T, first column corrected
      __syncwarp();
      for (int idx = lane; idx < ldab * nn; idx += 32) AB[idx] = 0.0;
      __syncwarp();
      for (int idx = lane; idx < 5 * nn; idx += 32) {
        int b = idx / nn, r = idx - b * nn;     // D(r, r+b-2) moves to row r+1
        int jc = r + b - 2, ir = r + 1;
        if (jc < 0 || jc >= nn || ir >= nn) continue;
        AB[(kv + ir - jc) + (size_t)jc * ldab] = D[b * nn + r];
      }
      __syncwarp();
      if (lane == 0) {
        AB[kv + 0] = 1.0;
        AB[kv + 1] = AB[kv + 1] + a.sp0;
        if (nn > 2) AB[kv + 2] = AB[kv + 2] - a.sp1;
        if (nn > 3) AB[kv + 3] = AB[kv + 3] + a.sp2;
      }
    }
    __syncwarp();

    // ---- zgbtf2 with the forward substitution of zgbtrs applied on the fly ----
    int ju = 0;
    for (int jj = 0; jj < nn; ++jj) {
      const int km = min(kl, nn - 1 - jj);
      double *cj = AB + (size_t)jj * ldab;
      double pv = (lane <= km) ? fabs(cj[kv + lane]) : -1.0;
      int jp = lane;
      warp_argmax_first(pv, jp);
      if (lane == 0) piv[jj] = (pv != 0.0) ? jp : 0;
      if (pv != 0.0) {
        ju = max(ju, min(jj + ku + jp, nn - 1));
        const int ncol = ju - jj + 1;
        if (jp != 0) {
          for (int cc = lane; cc < ncol; cc += 32) {
            double *p1 = AB + (size_t)(jj + cc) * ldab + (kv + jp - cc);
            double *p2 = AB + (size_t)(jj + cc) * ldab + (kv - cc);
            double t = *p1;
            *p1 = *p2;
            *p2 = t;
          }
          if (lane == 0) {
            double tr = rhs[2 * (jj + jp)], ti = rhs[2 * (jj + jp) + 1];
            rhs[2 * (jj + jp)] = rhs[2 * jj];
            rhs[2 * (jj + jp) + 1] = rhs[2 * jj + 1];
            rhs[2 * jj] = tr;
            rhs[2 * jj + 1] = ti;
          }
          __syncwarp();
        }
        if (km > 0) {
          const double rinv = 1.0 / cj[kv];
          __syncwarp();
          if (lane >= 1 && lane <= km) cj[kv + lane] = cj[kv + lane] * rinv;
          __syncwarp();
          const int nupd = km * (ncol - 1);
          for (int idx = lane; idx < nupd; idx += 32) {
            int r = 1 + idx % km, cc = 1 + idx / km;
            double *col2 = AB + (size_t)(jj + cc) * ldab;
            double temp = -col2[kv - cc];
            col2[kv + r - cc] = col2[kv + r - cc] + cj[kv + r] * temp;
          }
          if (lane >= 1 && lane <= km) {
            double l = cj[kv + lane];
            double tr = -rhs[2 * jj], ti = -rhs[2 * jj + 1];
            rhs[2 * (jj + lane)] = rhs[2 * (jj + lane)] + l * tr;
            rhs[2 * (jj + lane) + 1] = rhs[2 * (jj + lane) + 1] + l * ti;
          }
          __syncwarp();
        }
      } else if (lane == 0) {
        atomicOr(a.flag + 1, 1);   // 'lurc: lu factorization resulted in failure'
      }
    }
    if (a.fac_ab) {
      // keep the factors: every later solve with this operator only runs the two substitutions
      __syncwarp();
      const long long c0 = a.fac_off[j] + (long long)(k - a.k0) * nn;   // first column of this system
      const int lu = kl + ku + 1;
      double *fu = a.fac_ab + c0 * lu, *fl = a.fac_ab + a.fac_ncols * lu + c0 * kl;
      for (int idx = lane; idx < lu * nn; idx += 32) fu[idx] = AB[(size_t)(idx / lu) * ldab + idx % lu];
      for (int idx = lane; idx < kl * nn; idx += 32) fl[idx] = AB[(size_t)(idx / kl) * ldab + lu + idx % kl];
      for (int idx = lane; idx < nn; idx += 32) a.fac_piv[c0 + idx] = (unsigned char)piv[idx];
    }
    // ---- ztbsv: upper, no transpose, non-unit, bandwidth kv ----
    for (int jj = nn - 1; jj >= 0; --jj) {
      const double *cj = AB + (size_t)jj * ldab;
      double xr = rhs[2 * jj], xi = rhs[2 * jj + 1];
      if (xr != 0.0 || xi != 0.0) {
        const double ujj = cj[kv];
        xr = xr / ujj;
        xi = xi / ujj;
        __syncwarp();
        if (lane == 0) {
          rhs[2 * jj] = xr;
          rhs[2 * jj + 1] = xi;
        }
        const int cnt = min(jj, kv);
        if (lane >= 1 && lane <= cnt) {
          int i = jj - lane;
          double u = cj[kv - lane];
          rhs[2 * i] = rhs[2 * i] - xr * u;
          rhs[2 * i + 1] = rhs[2 * i + 1] - xi * u;
        }
      }
      __syncwarp();
    }
    for (int i = lane; i < nn; i += 32) col[i] = make_double2(rhs[2 * i], rhs[2 * i + 1]);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// solves with cached factors: zgbtrs only (forward sweep with the stored multipliers and pivots, then ztbsv),
// the same operations in the same order as the factor-and-solve kernel above, so results are bit-identical.
// One warp per (m,k) system; the right-hand side lives in shared memory, the factors stream from HBM exactly
// once per solve (they are the only traffic that matters: ldab*8 bytes per row against 16 bytes of field data).
// ---------------------------------------------------------------------------------------------
#define CSOLVE_WARPS 8

// MIRROR: the launch also serves the mirrored planes (mirror_mode 1 or 2); compiled separately so that the plain
// substitution keeps its code (the pass loop cost it 25 % at 256^3 when both lived in one kernel)
// (80 registers, three CTAs per SM.  Forcing four -- 64 registers, 68 bytes of spills -- made the 256^3 ihelmp solve slower:
// 1.46 vs 1.20 ms.)
template <bool MIRROR>
__global__ void __launch_bounds__(CSOLVE_WARPS * 32) band_solve_cached_kernel(SolveArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nsys = a.npl * a.nk;
  const int kl = a.kl, ku = a.ku, kv = kl + ku, ldab = 2 * kl + ku + 1;
  double *rhs = reinterpret_cast<double *>(smraw) + (size_t)warp * 2 * a.nnmax;
  for (int sys = blockIdx.x * CSOLVE_WARPS + warp; sys < nsys; sys += gridDim.x * CSOLVE_WARPS) {
    // consecutive systems of a block share the column j (same nn, neighbouring factor storage)
    const int j = sys / a.nk;
    const int kf = a.k0 + sys % a.nk;     // plane whose factors are used
    const int mglob = a.m0 + j * a.ms;
    const int nn = nn_of(mglob, a.nrc, a.npc);
    if (nn < 1) continue;
    const long long c0 = a.fac_off[j] + (long long)(kf - a.k0) * nn;
   for (int pass = 0; pass < (MIRROR ? 2 : 1); ++pass) {
    // pass 1: the plane nz - kf has the same operator (ak^2) and therefore the same factors; when the factor set
    // fits in L2 the second substitution reads them from there instead of HBM
    if (MIRROR && pass == 0 && a.mirror_mode == 2) continue;
    const int k = (!MIRROR || pass == 0) ? kf : a.mirror_nz - kf;
    if (MIRROR && pass == 1 && (kf == 0 || k < a.mirror_lo || k >= a.mirror_nz)) break;
    cplx *col = a.e + ((size_t)k * a.npl + j) * a.nrl;
    const bool special = (pass == 0 && a.special00 && mglob == 0 && k == 0);
    const double *__restrict__ FU = a.fac_ab + c0 * (kv + 1);
    const double *__restrict__ FL = a.fac_ab + a.fac_ncols * (kv + 1) + c0 * kl;
    const unsigned char *__restrict__ piv = a.fac_piv + c0;
    for (int i = lane; i < nn; i += 32) {
      cplx v;
      if (special && a.special00 == 2)
        v = (i == 0) ? make_double2(a.preln_rhs, 0.0) : col[i - 1];
      else
        v = col[i];
      rhs[2 * i] = v.x;
      rhs[2 * i + 1] = v.y;
    }
    __syncwarp();
    // ---- forward sweep (zgbtrs, external/lapack/SRC/zgbtrs.f:205-232) ----
    // The factors do not depend on the right-hand side, so they are fetched 8 columns ahead of the dependent
    // chain: lane = 8 u + i holds the multiplier L(i+1) of column jb + 4 h + u (kl <= 8), lanes 0..7 the pivots.
    {
      double lcur[2], lnxt[2] = {0.0, 0.0};
      int pcur, pnxt = 0;
      auto loadL = [&](int jb, double(&l)[2], int &pv) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int cc = jb + 4 * h + (lane >> 3);
          l[h] = (cc < nn && (lane & 7) < kl) ? __ldg(&FL[(size_t)cc * kl + (lane & 7)]) : 0.0;
        }
        pv = (lane < 8 && jb + lane < nn) ? (int)__ldg(&piv[jb + lane]) : 0;
      };
      loadL(0, lcur, pcur);
      for (int jb = 0; jb < nn; jb += 8) {
        if (jb + 8 < nn) loadL(jb + 8, lnxt, pnxt);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int jj = jb + u;
          if (jj >= nn) break;
          const int km = min(kl, nn - 1 - jj);
          const int jp = __shfl_sync(0xffffffffu, pcur, u);
          const double l = __shfl_sync(0xffffffffu, lcur[u >> 2], 8 * (u & 3) + ((lane - 1) & 7));   // lane i gets L(i)
          if (jp != 0) {
            if (lane == 0) {
              double tr = rhs[2 * (jj + jp)], ti = rhs[2 * (jj + jp) + 1];
              rhs[2 * (jj + jp)] = rhs[2 * jj];
              rhs[2 * (jj + jp) + 1] = rhs[2 * jj + 1];
              rhs[2 * jj] = tr;
              rhs[2 * jj + 1] = ti;
            }
            __syncwarp();
          }
          if (km > 0) {
            if (lane >= 1 && lane <= km) {
              double tr = -rhs[2 * jj], ti = -rhs[2 * jj + 1];
              rhs[2 * (jj + lane)] = rhs[2 * (jj + lane)] + l * tr;
              rhs[2 * (jj + lane) + 1] = rhs[2 * (jj + lane) + 1] + l * ti;
            }
            __syncwarp();
          }
        }
        lcur[0] = lnxt[0];
        lcur[1] = lnxt[1];
        pcur = pnxt;
      }
    }
    // ---- ztbsv: upper, no transpose, non-unit, bandwidth kv ----
    // lane t holds U(jj - t, jj) = AB[jj*ldab + kv - t] of 8 consecutive steps, fetched 8 steps ahead
    {
      double ucur[8], unxt[8];
      auto loadU = [&](int jtop, double(&u)[8]) {
#pragma unroll
        for (int sidx = 0; sidx < 8; ++sidx) {
          const int cc = jtop - sidx;
          u[sidx] = (cc >= 0 && lane <= kv && lane <= cc) ? __ldg(&FU[(size_t)cc * (kv + 1) + kv - lane]) : 0.0;
        }
      };
      loadU(nn - 1, ucur);
      for (int jt = nn - 1; jt >= 0; jt -= 8) {
#pragma unroll
        for (int sidx = 0; sidx < 8; ++sidx) unxt[sidx] = 0.0;
        if (jt - 8 >= 0) loadU(jt - 8, unxt);
#pragma unroll
        for (int sidx = 0; sidx < 8; ++sidx) {
          const int jj = jt - sidx;
          if (jj < 0) break;
          const double u = ucur[sidx];
          double xr = rhs[2 * jj], xi = rhs[2 * jj + 1];
          if (xr != 0.0 || xi != 0.0) {
            const double ujj = __shfl_sync(0xffffffffu, u, 0);
            xr = xr / ujj;
            xi = xi / ujj;
            __syncwarp();
            if (lane == 0) {
              rhs[2 * jj] = xr;
              rhs[2 * jj + 1] = xi;
            }
            const int cnt = min(jj, kv);
            if (lane >= 1 && lane <= cnt) {
              int i = jj - lane;
              rhs[2 * i] = rhs[2 * i] - xr * u;
              rhs[2 * i + 1] = rhs[2 * i + 1] - xi * u;
            }
          }
          __syncwarp();
        }
#pragma unroll
        for (int sidx = 0; sidx < 8; ++sidx) ucur[sidx] = unxt[sidx];
      }
    }
    for (int i = lane; i < nn; i += 32) col[i] = make_double2(rhs[2 * i], rhs[2 * i + 1]);
    __syncwarp();
   }   // pass
  }
}

// Narrow bands (ihelm, idel2: kl <= 3, kl + ku <= 5): the wide kernel above keeps 3 of 32 lanes busy in the forward
// sweep and 5 in the back substitution, and its throughput is set by the latency of one dependent column step per
// warp.  Here a warp carries FOUR systems, eight lanes each: lane g of a group owns the multiplier L(g) (forward) /
// the entry U(jj - g, jj) (backward) of its system, fetched 8 columns ahead of the dependent chain; same operations in
// the same order per system as the wide kernel (and as zgbtrs / ztbsv), so results are bit-identical.
// (Template parameter: lanes per system.  A 16-lane form for the hyperviscous operators, kl = ku <= 8 with lane 0 also
// carrying the entry at distance 16 of the back substitution, was measured slower than the wide kernel.)
#define NSOLVE_WARPS 4
template <bool MIRROR, int NSOLVE_GROUP>
__global__ void __launch_bounds__(NSOLVE_WARPS * 32) band_solve_cached_narrow_kernel(SolveArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int SPW = 32 / NSOLVE_GROUP;            // systems per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / NSOLVE_GROUP, gl = lane % NSOLVE_GROUP, gbase = grp * NSOLVE_GROUP;
  const int nsys = a.npl * a.nk;
  const int kl = a.kl, ku = a.ku, kv = kl + ku, ldab = 2 * kl + ku + 1;
  double *rhs = reinterpret_cast<double *>(smraw) + (size_t)(warp * SPW + grp) * 2 * a.nnmax;
  const int per_block = NSOLVE_WARPS * SPW;
  for (int sys0 = blockIdx.x * per_block + warp * SPW; sys0 < nsys; sys0 += gridDim.x * per_block) {
    const int sys = sys0 + grp;
    const bool have = sys < nsys;
    const int j = have ? sys / a.nk : 0;
    const int kf = a.k0 + (have ? sys % a.nk : 0);
    const int mglob = a.m0 + j * a.ms;
    const int nn = have ? nn_of(mglob, a.nrc, a.npc) : 0;
    // loop bounds are warp-uniform: the longest system of the warp (its groups share a column except at column ends)
    int nnw = nn;
#pragma unroll
    for (int off = NSOLVE_GROUP; off < 32; off <<= 1) nnw = max(nnw, __shfl_xor_sync(0xffffffffu, nnw, off));
    const long long c0 = have ? a.fac_off[j] + (long long)(kf - a.k0) * nn : 0;
    const double *__restrict__ FU = a.fac_ab + c0 * (kv + 1);
    const double *__restrict__ FL = a.fac_ab + a.fac_ncols * (kv + 1) + c0 * kl;
    const unsigned char *__restrict__ piv = a.fac_piv + c0;
    for (int pass = 0; pass < (MIRROR ? 2 : 1); ++pass) {
      if (MIRROR && pass == 0 && a.mirror_mode == 2) continue;
      const int k = (!MIRROR || pass == 0) ? kf : a.mirror_nz - kf;
      // a group whose mirrored plane does not exist sits this pass out (nnp = 0); the warp stays together
      const bool live = nn > 0 && !(MIRROR && pass == 1 && (kf == 0 || k < a.mirror_lo || k >= a.mirror_nz));
      const int nnp = live ? nn : 0;
      cplx *col = a.e + ((size_t)(live ? k : 0) * a.npl + j) * a.nrl;
      const bool special = (pass == 0 && a.special00 && mglob == 0 && k == 0 && live);
      for (int i = gl; i < nnp; i += NSOLVE_GROUP) {
        cplx v;
        if (special && a.special00 == 2)
          v = (i == 0) ? make_double2(a.preln_rhs, 0.0) : col[i - 1];
        else
          v = col[i];
        rhs[2 * i] = v.x;
        rhs[2 * i + 1] = v.y;
      }
      __syncwarp();
      // ---- forward sweep (zgbtrs, external/lapack/SRC/zgbtrs.f:205-232) ----
      {
        double lcur[8], lnxt[8];
        int pcur[8], pnxt[8];
        auto loadL = [&](int jb, double(&l)[8], int(&pv)[8]) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int cc = jb + u;
            l[u] = (cc < nnp && gl >= 1 && gl <= kl) ? __ldg(&FL[(size_t)cc * kl + gl - 1]) : 0.0;
            pv[u] = (cc < nnp && gl == 0) ? (int)__ldg(&piv[cc]) : 0;
          }
        };
        loadL(0, lcur, pcur);
        for (int jb = 0; jb < nnw; jb += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            lnxt[u] = 0.0;
            pnxt[u] = 0;
          }
          if (jb + 8 < nnw) loadL(jb + 8, lnxt, pnxt);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int jj = jb + u;
            if (jj >= nnw) break;
            const bool on = jj < nnp;
            const int km = on ? min(kl, nnp - 1 - jj) : 0;
            const int jp = __shfl_sync(0xffffffffu, pcur[u], gbase);
            if (on && jp != 0 && gl == 0) {
              double tr = rhs[2 * (jj + jp)], ti = rhs[2 * (jj + jp) + 1];
              rhs[2 * (jj + jp)] = rhs[2 * jj];
              rhs[2 * (jj + jp) + 1] = rhs[2 * jj + 1];
              rhs[2 * jj] = tr;
              rhs[2 * jj + 1] = ti;
            }
            __syncwarp();
            if (gl >= 1 && gl <= km) {
              const double l = lcur[u];
              double tr = -rhs[2 * jj], ti = -rhs[2 * jj + 1];
              rhs[2 * (jj + gl)] = rhs[2 * (jj + gl)] + l * tr;
              rhs[2 * (jj + gl) + 1] = rhs[2 * (jj + gl) + 1] + l * ti;
            }
            __syncwarp();
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            lcur[u] = lnxt[u];
            pcur[u] = pnxt[u];
          }
        }
      }
      // ---- ztbsv: upper, no transpose, non-unit, bandwidth kv; the systems of a warp run bottom-aligned ----
      {
        double ucur[8], unxt[8];
        double vcur[8], vnxt[8];                     // entry at distance gl + G (only when kv >= G: G = 16, kv = 16)
        const int g2 = gl + NSOLVE_GROUP;
        const bool second = NSOLVE_GROUP == 16 && g2 <= kv;
        auto loadU = [&](int jtop, double(&u)[8], double(&v)[8]) {
#pragma unroll
          for (int sidx = 0; sidx < 8; ++sidx) {
            const int cc = jtop - sidx;
            u[sidx] = (cc >= 0 && cc < nnp && gl <= kv && gl <= cc) ? __ldg(&FU[(size_t)cc * (kv + 1) + kv - gl]) : 0.0;
            if (NSOLVE_GROUP == 16)
              v[sidx] = (second && cc >= 0 && cc < nnp && g2 <= cc) ? __ldg(&FU[(size_t)cc * (kv + 1) + kv - g2]) : 0.0;
          }
        };
        loadU(nnw - 1, ucur, vcur);
        for (int jt = nnw - 1; jt >= 0; jt -= 8) {
#pragma unroll
          for (int sidx = 0; sidx < 8; ++sidx) unxt[sidx] = vnxt[sidx] = 0.0;
          if (jt - 8 >= 0) loadU(jt - 8, unxt, vnxt);
#pragma unroll
          for (int sidx = 0; sidx < 8; ++sidx) {
            const int jj = jt - sidx;
            if (jj < 0) break;
            const double u = ucur[sidx];
            const double ujj = __shfl_sync(0xffffffffu, u, gbase);
            double xr = 0.0, xi = 0.0;
            if (jj < nnp) {
              xr = rhs[2 * jj];
              xi = rhs[2 * jj + 1];
            }
            const bool act = jj < nnp && (xr != 0.0 || xi != 0.0);
            if (act) {
              xr = xr / ujj;
              xi = xi / ujj;
            }
            __syncwarp();
            if (act) {
              if (gl == 0) {
                rhs[2 * jj] = xr;
                rhs[2 * jj + 1] = xi;
              }
              const int cnt = min(jj, kv);
              if (gl >= 1 && gl <= cnt) {
                const int i = jj - gl;
                rhs[2 * i] = rhs[2 * i] - xr * u;
                rhs[2 * i + 1] = rhs[2 * i + 1] - xi * u;
              }
              if (NSOLVE_GROUP == 16 && second && g2 <= cnt) {
                const int i = jj - g2;
                rhs[2 * i] = rhs[2 * i] - xr * vcur[sidx];
                rhs[2 * i + 1] = rhs[2 * i + 1] - xi * vcur[sidx];
              }
            }
            __syncwarp();
          }
#pragma unroll
          for (int sidx = 0; sidx < 8; ++sidx) {
            ucur[sidx] = unxt[sidx];
            vcur[sidx] = vnxt[sidx];
          }
        }
      }
      for (int i = gl; i < nnp; i += NSOLVE_GROUP) col[i] = make_double2(rhs[2 * i], rhs[2 * i + 1]);
      __syncwarp();
    }   // pass
  }
}

// ---- factor cache ---------------------------------------------------------------------------------------------
struct FactorEntry {
  SolveArgs key;            // operator identity (pointer fields zeroed)
  double *ab = nullptr;
  unsigned char *piv = nullptr;
  long long *off = nullptr;
  long long ncols = 0;
  size_t bytes = 0;
  unsigned long long stamp = 0;
};
static std::vector<FactorEntry> g_fcache;
static unsigned long long g_fstamp = 0;
static int g_fcache_on = 1;

static void free_entry(FactorEntry &e) {
  if (e.ab) cudaFree(e.ab);
  if (e.piv) cudaFree(e.piv);
  if (e.off) cudaFree(e.off);
  e = FactorEntry();
}
void band_solve_cache_clear() {
  for (auto &e : g_fcache) free_entry(e);
  g_fcache.clear();
}
void band_solve_cache_enable(int on) {
  g_fcache_on = on;
  if (!on) band_solve_cache_clear();
}

static SolveArgs key_of(const SolveArgs &a) {
  SolveArgs k;
  memset(&k, 0, sizeof(k));   // padding bytes too: keys are compared with memcmp
  k.nrl = a.nrl; k.npl = a.npl; k.m0 = a.m0; k.ms = a.ms; k.k0 = a.k0; k.nk = a.nk; k.ne = a.ne;
  k.nrc = a.nrc; k.npc = a.npc; k.nnmax = a.nnmax; k.kl = a.kl; k.ku = a.ku; k.power = a.power;
  k.add_alpha = a.add_alpha; k.alpha = a.alpha; k.beta = a.beta; k.special00 = a.special00;
  k.sp0 = a.sp0; k.sp1 = a.sp1; k.sp2 = a.sp2;   // preln_rhs only enters the right-hand side
  return k;
}

static bool factor_cached(const SolveArgs &a) {
  if (!g_fcache_on) return false;
  SolveArgs key = key_of(a);
  for (auto &e : g_fcache)
    if (memcmp(&e.key, &key, sizeof(SolveArgs)) == 0) return true;
  return false;
}

// Planes [0, n1) and [lo, nzl) of one operator.  The operator of plane k only contains ak(k)^2 = ak(nz-k)^2, so the
// second range can be served with the factors of the first: half the factor memory, and -- when the factor set is
// small enough to survive in L2 between the two substitutions of a warp -- less HBM traffic (128^3: 0.43 -> 0.31 ms
// per step).  For large factor sets the mirrored substitutions miss L2 and only halve the parallelism (256^3:
// 1.77 -> 2.14 ms; two right-hand sides through one sweep: 1.90 ms), so there the ranges stay two launches with
// their own factors.
int launch_band_solve_ranges(SolveArgs base, int n1, int kl_first, int lo, int nzl, cudaStream_t st) {
  base.mirror_mode = 0;
  base.mirror_lo = base.mirror_nz = 0;
  SolveArgs a = base;
  a.k0 = 0;
  a.nk = n1;
  a.kl = kl_first;
  SolveArgs b = base;
  b.k0 = lo;
  b.nk = nzl - lo;
  b.special00 = 0;
  const bool have_a = n1 > 0, have_b = lo < nzl;
  size_t cols = 0;   // factor columns of the first range
  for (int j = 0; j < a.npl; ++j) {
    const int m = a.m0 + j * a.ms;
    cols += (size_t)((m < a.npc) ? std::max(std::min(a.nrc, a.nrc - m), 0) : 0) * (size_t)std::max(n1, 0);
  }
  const size_t fbytes = cols * (size_t)(2 * a.kl + a.ku + 1) * sizeof(double);
  // every plane k of the second range has its mirror nz - k inside the first one, same band widths
  const bool mirrored = have_a && have_b && g_fcache_on && a.kl <= 8 && a.kl == b.kl && lo >= 1 && nzl - lo <= n1 - 1 &&
                        fbytes <= ((size_t)96 << 20);
  if (mirrored) {
    a.mirror_lo = lo;
    a.mirror_nz = nzl;
    if (factor_cached(a)) {
      a.mirror_mode = 1;                       // one launch: (j,k) then (j, nz-k) by the same warp
      return launch_band_solve(a, st);
    }
    MLEGS_TRY(launch_band_solve(a, st));       // factors (and keeps them if memory allows), solves the first range
    if (factor_cached(a)) {
      a.mirror_mode = 2;                       // second range from the factors just stored
      return launch_band_solve(a, st);
    }
    return launch_band_solve(b, st);
  }
  if (have_a) MLEGS_TRY(launch_band_solve(a, st));
  if (have_b) MLEGS_TRY(launch_band_solve(b, st));
  return MLEGS_OK;
}

int launch_band_solve(SolveArgs a, cudaStream_t st) {
  Context &c = ctx();
  if (a.npl <= 0 || a.nk <= 0) return MLEGS_OK;
  const int ldab = 2 * a.kl + a.ku + 1;
  a.flag = c.d_flag;
  a.fac_ab = nullptr;
  a.fac_piv = nullptr;
  a.fac_off = nullptr;
  a.fac_ncols = 0;

  // ---- cached factors? ----
  FactorEntry *hit = nullptr;
  SolveArgs key = key_of(a);
  if (g_fcache_on) {
    for (auto &e : g_fcache)
      if (memcmp(&e.key, &key, sizeof(SolveArgs)) == 0) hit = &e;
  }
  if (hit) {
    hit->stamp = ++g_fstamp;
    a.fac_ab = hit->ab;
    a.fac_piv = hit->piv;
    a.fac_off = hit->off;
    a.fac_ncols = hit->ncols;
    const int nsys = a.npl * a.nk;
    size_t smem = (size_t)CSOLVE_WARPS * 2 * a.nnmax * sizeof(double);
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
      CUDA_TRY(cudaFuncSetAttribute(band_solve_cached_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CUDA_TRY(cudaFuncSetAttribute(band_solve_cached_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = smem;
    }
    int blocks = (nsys + CSOLVE_WARPS - 1) / CSOLVE_WARPS;
    // algorithmic bytes: the factor stream (ldab doubles + one pivot byte per matrix column, read once -- the
    // mirrored plane re-reads it from L2) + one read and one write of every right-hand side
    double syscols = 0.0;
    for (int j = 0; j < a.npl; ++j) {
      const int m = a.m0 + j * a.ms;
      syscols += (double)((m < a.npc) ? std::max(std::min(a.nrc, a.nrc - m), 0) : 0) * a.nk;
    }
    const double rhs_cols = a.mirror_mode == 1 ? 2.0 * syscols : syscols;
    prof_begin(a.power > 2 ? "ihelmp_solve_cached" : "band_solve_cached", st,
               syscols * (ldab * 8.0 + 1.0) + 32.0 * rhs_cols);
    static const bool no_narrow = getenv("MLEGS_SOLVE_WIDE") != nullptr;   // A/B timing
    // Several systems per warp pay when there are more systems than the wide kernel keeps resident (256^3: ihelm / idel2
    // 0.56 -> 0.45 ms); with fewer, the latency of one system decides and the wide kernel is faster (128^3: 0.095 vs
    // 0.13 ms).  The 16-lane form for the hyperviscous operators lost at every size (256^3: 1.30 vs 1.21 ms) and is not
    // dispatched.
    const int gsz = (a.kl <= 3 && a.kl + a.ku <= 5 && nsys >= 16384) ? 8 : 0;
    if (gsz && !no_narrow) {
      const int spw = 32 / gsz;
      const size_t nsmem = (size_t)NSOLVE_WARPS * spw * 2 * a.nnmax * sizeof(double);
      static size_t nattr = 0;
      if (nsmem > 48 * 1024 && nsmem > nattr) {
        CUDA_TRY(cudaFuncSetAttribute(band_solve_cached_narrow_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nsmem));
        CUDA_TRY(cudaFuncSetAttribute(band_solve_cached_narrow_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nsmem));
        nattr = nsmem;
      }
      const int nblocks = (nsys + NSOLVE_WARPS * spw - 1) / (NSOLVE_WARPS * spw);
      if (a.mirror_mode != 0)
        band_solve_cached_narrow_kernel<true, 8><<<nblocks, NSOLVE_WARPS * 32, nsmem, st>>>(a);
      else
        band_solve_cached_narrow_kernel<false, 8><<<nblocks, NSOLVE_WARPS * 32, nsmem, st>>>(a);
    } else if (a.mirror_mode != 0)
      band_solve_cached_kernel<true><<<blocks, CSOLVE_WARPS * 32, smem, st>>>(a);
    else
      band_solve_cached_kernel<false><<<blocks, CSOLVE_WARPS * 32, smem, st>>>(a);
    prof_end(st);
    KERNEL_CHECK();
    return MLEGS_OK;
  }

  // ---- first solve with this operator: factor (and keep the factors if memory allows) ----
  if (g_fcache_on && a.kl <= 8) {
    std::vector<long long> off(a.npl + 1, 0);
    for (int j = 0; j < a.npl; ++j) {
      int m = a.m0 + j * a.ms;
      int nn = (m < a.npc) ? std::max(std::min(a.nrc, a.nrc - m), 0) : 0;
      off[j + 1] = off[j] + (long long)nn * a.nk;
    }
    const size_t ncols = (size_t)off[a.npl];
    const size_t need = ncols * ldab * sizeof(double) + ncols + off.size() * sizeof(long long);
    size_t freeb = 0, totalb = 0;
    cudaMemGetInfo(&freeb, &totalb);
    // The cache is keyed on the exact (alpha, beta): a caller that varies dt adds a factor set per step.  Bound it by
    // entry count and by a byte budget (a third of the device memory) as well as by what is free right now, evicting
    // least recently used entries first.
    const size_t kMaxEntries = 12;
    size_t held = 0;
    for (auto &e : g_fcache) held += e.bytes;
    while (!g_fcache.empty() && (need > freeb / 2 || g_fcache.size() >= kMaxEntries || held + need > totalb / 3)) {
      size_t lru = 0;
      for (size_t i = 1; i < g_fcache.size(); ++i)
        if (g_fcache[i].stamp < g_fcache[lru].stamp) lru = i;
      CUDA_TRY(cudaStreamSynchronize(st));
      held -= g_fcache[lru].bytes;
      free_entry(g_fcache[lru]);
      g_fcache.erase(g_fcache.begin() + lru);
      cudaMemGetInfo(&freeb, &totalb);
    }
    if (ncols > 0 && need <= freeb / 2) {
      FactorEntry e;
      e.key = key;
      e.bytes = need;
      e.stamp = ++g_fstamp;
      bool ok = cudaMalloc((void **)&e.ab, ncols * ldab * sizeof(double)) == cudaSuccess &&
                cudaMalloc((void **)&e.piv, ncols) == cudaSuccess &&
                cudaMalloc((void **)&e.off, off.size() * sizeof(long long)) == cudaSuccess;
      if (ok) {
        CUDA_TRY(cudaMemcpyAsync(e.off, off.data(), off.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));   // `off` goes out of scope
        e.ncols = (long long)ncols;
        g_fcache.push_back(e);
        a.fac_ab = e.ab;
        a.fac_piv = e.piv;
        a.fac_off = e.off;
        a.fac_ncols = e.ncols;
      } else {
        cudaGetLastError();
        free_entry(e);
      }
    }
  }

  a.ws_doubles = (size_t)(5 + (a.power > 2 ? 17 : 0) + ldab + 2) * a.nnmax + (a.nnmax + 1) / 2;   // + pivots (ints)
  // Pb (used from power 6 on) aliases AB and holds at most 13 bands stored with offset 8 -> 15 nn doubles
  if (a.power > 4 && ldab < 15) return fail(MLEGS_E_ARG, "band_solve: internal workspace aliasing violated");
  size_t smem = a.ws_doubles * sizeof(double) * SOLVE_WARPS;
  const int nsys = a.npl * a.nk;
  int blocks = (nsys + SOLVE_WARPS - 1) / SOLVE_WARPS;
  if (smem <= 200 * 1024) {
    a.ws_global = nullptr;
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
      CUDA_TRY(cudaFuncSetAttribute(band_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = smem;
    }
  } else {
    // large systems: per-warp workspace in global memory (L2 resident), persistent grid
    smem = 0;
    blocks = std::min(blocks, 148 * 8);
    size_t need = (size_t)blocks * SOLVE_WARPS * a.ws_doubles * sizeof(double);
    if (need > c.solve_ws_bytes) {
      if (c.d_solve_ws) CUDA_TRY(cudaFree(c.d_solve_ws));
      CUDA_TRY(cudaMalloc(&c.d_solve_ws, need));
      c.solve_ws_bytes = need;
    }
    a.ws_global = (double *)c.d_solve_ws;
  }
  prof_begin(a.power > 2 ? "ihelmp_solve" : "band_solve", st);
  band_solve_kernel<<<blocks, SOLVE_WARPS * 32, smem, st>>>(a);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

}  // namespace mlegs
