// Fused pointwise kernels: truncation masks, the SVV filter, far-field values, the diagonal
// (1-x)^-2 del^2_perp operators, the nonlinear cross product and the integrators' axpys.
// Reference: /root/reference/src/submodules/mlegs_scalar_ops.f90:6-155, 237-416, 1264-1306.
// All HBM-bound; compiled with -fmad=false so that every product/sum rounds like the reference's
// (non-fused) Fortran expressions.
#include <algorithm>
#include <cmath>

#include "kernels.h"

namespace mlegs {

#define EW_THREADS 256

static inline unsigned ew_grid(size_t n) {
  size_t g = (n + EW_THREADS - 1) / EW_THREADS;
  const size_t cap = 148 * 16;
  return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

// ---------------------------------------------------------------------------------------------
// masks: chop (ops:6-41) and dealias (ops:43-70)
// ---------------------------------------------------------------------------------------------

// One warp per (column, plane) line: the line is zeroed as a whole (column beyond the azimuthal cut, plane inside the
// axial cut) or from its truncation row on; retained entries are never read or written.
__global__ void __launch_bounds__(EW_THREADS) mask_kernel(MaskArgs a) {
  const int lane = threadIdx.x & 31;
  const long long nlines = (long long)a.npl * a.nzl;
  const long long wstride = (long long)gridDim.x * (EW_THREADS / 32);
  for (long long line = (long long)blockIdx.x * (EW_THREADS / 32) + (threadIdx.x >> 5); line < nlines; line += wstride) {
    const int j = (int)(line % a.npl), k = (int)(line / a.npl);
    const int m = a.m0 + j * a.ms;
    int first = a.nrl;                                 // first zeroed row of this line
    if (m >= a.col_cut || (k >= a.kz_lo && k < a.kz_hi)) {
      first = 0;
    } else if (a.row_mode) {
      int nn = 0;
      if (m < a.npc_rows) {
        nn = min(a.nrc, a.nrc - m);
        nn = nn > 0 ? nn : 0;
      }
      first = min(max(nn - a.r0, 0), a.nrl);
    }
    cplx *col = a.e + (size_t)line * a.nrl;
    for (int i = first + lane; i < a.nrl; i += 32) col[i] = make_double2(0.0, 0.0);
  }
}

int launch_mask(const MaskArgs &a, cudaStream_t st) {
  size_t n = (size_t)a.nrl * a.npl * a.nzl;
  if (!n) return MLEGS_OK;
  // write-only: 16 bytes per zeroed element
  double zeroed = 0.0;
  {
    int kzm = 0;
    for (int k = 0; k < a.nzl; ++k) kzm += (k >= a.kz_lo && k < a.kz_hi) ? 1 : 0;
    for (int j = 0; j < a.npl; ++j) {
      const int m = a.m0 + j * a.ms;
      if (m >= a.col_cut) {
        zeroed += (double)a.nrl * a.nzl;
        continue;
      }
      int rz = 0;
      if (a.row_mode) {
        int nn = m < a.npc_rows ? std::max(std::min(a.nrc, a.nrc - m), 0) : 0;
        rz = a.nrl - std::min(std::max(nn - a.r0, 0), a.nrl);
      }
      zeroed += (double)kzm * a.nrl + (double)(a.nzl - kzm) * rz;
    }
  }
  prof_begin("mask", st, 16.0 * zeroed);
  mask_kernel<<<ew_grid((size_t)a.npl * a.nzl * 32), EW_THREADS, 0, st>>>(a);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// ---------------------------------------------------------------------------------------------
// SVV filter, ops:72-155
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ double svv_q(const SvvArgs &a, int i, int j, int k) {
  double q_r = fmin(1.0, (double)max(a.r0 + i, 0) / a.qr_den);
  double q_p = fmin(1.0, (double)abs(a.m0 + j * a.ms) / a.qp_den);
  double kv = (k < a.nak) ? fabs(a.ak[k]) : 0.0;
  double q_z = fmin(1.0, kv / a.kmax);
  return fmin(1.0, fmax(fmax(q_r, q_p), q_z));
}

#define SVV_BLOCKS 1024
__global__ void svv_energy_kernel(SvvArgs a, double *partial /* [2][SVV_BLOCKS] */) {
  __shared__ double s_tot[EW_THREADS], s_tail[EW_THREADS];
  const size_t n = (size_t)a.nrl * a.npl * a.nzl;
  double tot = 0.0, tail = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    int i = (int)(idx % a.nrl);
    size_t t = idx / a.nrl;
    int j = (int)(t % a.npl);
    int k = (int)(t / a.npl);
    cplx v = a.e[idx];
    double e2 = v.x * v.x + v.y * v.y;
    tot += e2;
    if (svv_q(a, i, j, k) >= a.cutoff) tail += e2;
  }
  s_tot[threadIdx.x] = tot;
  s_tail[threadIdx.x] = tail;
  __syncthreads();
  for (int s = EW_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      s_tot[threadIdx.x] += s_tot[threadIdx.x + s];
      s_tail[threadIdx.x] += s_tail[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = s_tot[0];
    partial[SVV_BLOCKS + blockIdx.x] = s_tail[0];
  }
}

__global__ void svv_reduce_kernel(const double *partial, int nblocks, double *out2) {
  __shared__ double s_tot[EW_THREADS], s_tail[EW_THREADS];
  double tot = 0.0, tail = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += EW_THREADS) {
    tot += partial[i];
    tail += partial[SVV_BLOCKS + i];
  }
  s_tot[threadIdx.x] = tot;
  s_tail[threadIdx.x] = tail;
  __syncthreads();
  for (int s = EW_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      s_tot[threadIdx.x] += s_tot[threadIdx.x + s];
      s_tail[threadIdx.x] += s_tail[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out2[0] = s_tot[0];
    out2[1] = s_tail[0];
  }
}

__global__ void svv_apply_kernel(SvvArgs a) {
  const size_t n = (size_t)a.nrl * a.npl * a.nzl;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    int i = (int)(idx % a.nrl);
    size_t t = idx / a.nrl;
    int j = (int)(t % a.npl);
    int k = (int)(t / a.npl);
    double q = svv_q(a, i, j, k);
    if (q > a.cutoff) {
      double shape = (q - a.cutoff) / (1.0 - a.cutoff);
      double s2 = shape * shape, s4 = s2 * s2;
      double factor = exp(-a.strength * (s4 * s4));
      cplx v = a.e[idx];
      a.e[idx] = make_double2(factor * v.x, factor * v.y);
    }
  }
}

int launch_svv_energy(const SvvArgs &a, double *d_partial, double *d_out2, cudaStream_t st) {
  size_t n = (size_t)a.nrl * a.npl * a.nzl;
  unsigned g = ew_grid(n);
  if (g > SVV_BLOCKS) g = SVV_BLOCKS;
  prof_begin("svv_energy", st, 16.0 * (double)n);
  svv_energy_kernel<<<g, EW_THREADS, 0, st>>>(a, d_partial);
  prof_end(st);
  KERNEL_CHECK();
  svv_reduce_kernel<<<1, EW_THREADS, 0, st>>>(d_partial, (int)g, d_out2);
  KERNEL_CHECK();
  return MLEGS_OK;
}

int launch_svv_apply(const SvvArgs &a, cudaStream_t st) {
  size_t n = (size_t)a.nrl * a.npl * a.nzl;
  prof_begin("svv_apply", st);
  svv_apply_kernel<<<ew_grid(n), EW_THREADS, 0, st>>>(a);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// ---------------------------------------------------------------------------------------------
// calcat0 / calcat1 / zeroat1, ops:237-325: calc(k) = sum_n e(n, m=0, k) * at(n)
// one CTA per axial mode; fixed-order tree reduction (deterministic)
// ---------------------------------------------------------------------------------------------
__global__ void calcat_kernel(cplx *e, int nrl, int npl, int nrows, const double *at, cplx *out, int subtract,
                              double at_first) {
  __shared__ double sr[EW_THREADS], si[EW_THREADS];
  const int k = blockIdx.x;
  cplx *col = e + (size_t)k * nrl * npl;   // local column 0 == global m 0 (caller guarantees)
  double ar = 0.0, ai = 0.0;
  for (int n = threadIdx.x; n < nrows; n += EW_THREADS) {
    cplx v = col[n];
    double w = at[n];
    ar += v.x * w;
    ai += v.y * w;
  }
  sr[threadIdx.x] = ar;
  si[threadIdx.x] = ai;
  __syncthreads();
  for (int s = EW_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sr[threadIdx.x] += sr[threadIdx.x + s];
      si[threadIdx.x] += si[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (out) out[k] = make_double2(sr[0], si[0]);
    if (subtract) {   // zeroat1: e(1,1,k) -= at1(k)/tfm%at1(1)
      cplx v = col[0];
      col[0] = make_double2(v.x - sr[0] / at_first, v.y - si[0] / at_first);
    }
  }
}

int launch_calcat(cplx *e, int nrl, int npl, int nzl, int nrows, const double *at, cplx *out, int subtract,
                  double at_first, cudaStream_t st) {
  prof_begin("calcat", st, 16.0 * (double)nrows * nzl);
  calcat_kernel<<<nzl, EW_THREADS, 0, st>>>(e, nrl, npl, nrows, at, out, subtract, at_first);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// ---------------------------------------------------------------------------------------------
// delsqp / idelsqp, ops:327-416: e(n,m,:) *= -n'(n'+1)/ell^2 (or its inverse), n' = m + n
// ---------------------------------------------------------------------------------------------
__global__ void delsqp_kernel(cplx *e, int nrl, int npl, int nzl, int m0, int ms, int nrc, int npc, double ell2, int inverse) {
  const size_t n = (size_t)nrl * npl * nzl;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    int i = (int)(idx % nrl);
    size_t t = idx / nrl;
    int j = (int)(t % npl);
    int m = m0 + j * ms;
    if (m >= npc) continue;
    int nn = min(nrc, nrc - m);
    if (i >= nn) continue;
    int nq = m + i;
    cplx v = e[idx];
    if (!inverse) {
      // -s%e*n*(n+1.D0)/(ell**2.D0)
      double fr = ((-v.x) * nq) * (nq + 1.0) / ell2;
      double fi = ((-v.y) * nq) * (nq + 1.0) / ell2;
      e[idx] = make_double2(fr, fi);
    } else if (nq == 0) {
      e[idx] = make_double2(0.0, 0.0);
    } else {
      // -s%e/n/(n+1.D0)*(ell**2.D0)
      double fr = (-v.x) / nq / (nq + 1.0) * ell2;
      double fi = (-v.y) / nq / (nq + 1.0) * ell2;
      e[idx] = make_double2(fr, fi);
    }
  }
}

int launch_delsqp(cplx *e, int nrl, int npl, int nzl, int m0, int ms, int nrc, int npc, double ell2, int inverse,
                  cudaStream_t st) {
  size_t n = (size_t)nrl * npl * nzl;
  prof_begin(inverse ? "idelsqp" : "delsqp", st, 32.0 * retained_elems(nrl, npl, nzl, 0, m0, nrc, npc, nzl, nzl, ms));
  delsqp_kernel<<<ew_grid(n), EW_THREADS, 0, st>>>(e, nrl, npl, nzl, m0, ms, nrc, npc, ell2, inverse);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// set / add a few individual entries (log-term corrections, e.g. ops:358-360, 453-456, 562-566)
__global__ void poke_kernel(cplx *e, PokeArgs p) {
  int t = threadIdx.x;
  if (t < p.n) {
    if (p.mode[t] == 0)
      e[p.off[t]] = make_double2(p.re[t], p.im[t]);
    else {
      cplx v = e[p.off[t]];
      e[p.off[t]] = make_double2(v.x + p.re[t], v.y + p.im[t]);
    }
  }
}
int launch_poke(cplx *e, const PokeArgs &p, cudaStream_t st) {
  if (p.n <= 0) return MLEGS_OK;
  poke_kernel<<<1, 32, 0, st>>>(e, p);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// zero a strided line: e[off + t*stride] = 0, t < n  (idelsqp: so%e(1,1,:) = 0, ops:408-410)
__global__ void zero_line_kernel(cplx *e, long long off, long long stride, int n) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
    e[off + (long long)t * stride] = make_double2(0.0, 0.0);
}
int launch_zero_line(cplx *e, long long off, long long stride, int n, cudaStream_t st) {
  zero_line_kernel<<<(n + 255) / 256, 256, 0, st>>>(e, off, stride, n);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// ---------------------------------------------------------------------------------------------
// vector product, ops:1264-1306: separate cross products on the Re (phi_{2j-1}) and Im (phi_{2j}) lanes
// ---------------------------------------------------------------------------------------------
__global__ void vecprod_kernel(cplx *vr, cplx *vp, cplx *vz, const cplx *ur, const cplx *up, const cplx *uz, int nrl,
                               int npl, int nzl, int r0, int nr, int nph, int nz) {
  const size_t n = (size_t)nrl * npl * nzl;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    int i = (int)(idx % nrl);
    size_t t = idx / nrl;
    int j = (int)(t % npl);
    int k = (int)(t / npl);
    cplx o1 = make_double2(0.0, 0.0), o2 = o1, o3 = o1;
    if (r0 + i < nr && j < nph && k < nz) {
      cplx a = vr[idx], b = vp[idx], c = vz[idx];
      cplx d = ur[idx], f = up[idx], g = uz[idx];
      o1 = make_double2(b.x * g.x - c.x * f.x, b.y * g.y - c.y * f.y);
      o2 = make_double2(c.x * d.x - a.x * g.x, c.y * d.y - a.y * g.y);
      o3 = make_double2(a.x * f.x - b.x * d.x, a.y * f.y - b.y * d.y);
    }
    vr[idx] = o1;
    vp[idx] = o2;
    vz[idx] = o3;
  }
}

int launch_vecprod(cplx *vr, cplx *vp, cplx *vz, const cplx *ur, const cplx *up, const cplx *uz, int nrl, int npl,
                   int nzl, int r0, int nr, int nph, int nz, cudaStream_t st) {
  size_t n = (size_t)nrl * npl * nzl;
  prof_begin("vecprod", st, 9.0 * 16.0 * (double)n);   // 6 fields read, 3 written
  vecprod_kernel<<<ew_grid(n), EW_THREADS, 0, st>>>(vr, vp, vz, ur, up, uz, nrl, npl, nzl, r0, nr, nph, nz);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// ---------------------------------------------------------------------------------------------
// linear combinations used by the integrators and the apps' whole-array statements
//   mode 0: y = a*x + b*y                    (apps/vortical_flow_3d.f90:136-137: 2*psi - psi_rich)
//   mode 1: y = y + a*x                      (vz%e = vz%e + uz%e, :379)
//   mode 2: out = s + dt*nl                  (febe,  ops:1169)
//   mode 3: out = s + dt*(1.5*nl - 0.5*nlp)  (abcn,  ops:1214)
//   mode 4: y = a*y                          (sh%e = a*sh%e, ops:1180)
//   mode 5: y = a*(y + b*(c*x))              (abcn: a*(sh + dt/2*(hv*svis)), ops:1229,1251)
//   mode 6: out = sp + beta*s2 + alpha*s     (helmp, ops:893)
//   mode 7: y = y + dt*(x1 + c*x2)           (fefe: s + dt*(nl + hv*svis), ops:1086-1089)
//   mode 8: y = y + a*(1.5*(x1 + c*x2) - 0.5*(x3 + d*x4))   (abab, ops:1147)
//   mode 9: y = a*((y + d*(1.5*x1 - 0.5*x2)) + b*(c*x3))   (abcn: modes 3 and 5 in one pass, ops:1214 + 1229/1251;
//           the inner sum is rounded to double exactly where the reference stores it in sh%e)
// ---------------------------------------------------------------------------------------------

__global__ void lincomb_kernel(LinArgs p) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < p.n; idx += (size_t)gridDim.x * blockDim.x) {
    cplx y = make_double2(0.0, 0.0), o;
    if (p.mode != 2 && p.mode != 3 && p.mode != 6) y = p.y[idx];   // out-of-place modes never read the destination
    switch (p.mode) {
      case 0: {
        cplx x = p.x1[idx];
        o = make_double2(p.a * x.x + p.b * y.x, p.a * x.y + p.b * y.y);
        break;
      }
      case 1: {
        cplx x = p.x1[idx];
        o = make_double2(y.x + p.a * x.x, y.y + p.a * x.y);
        break;
      }
      case 2: {
        cplx s = p.x1[idx], nl = p.x2[idx];
        o = make_double2(s.x + p.a * nl.x, s.y + p.a * nl.y);
        break;
      }
      case 3: {
        cplx s = p.x1[idx], nl = p.x2[idx], nlp = p.x3[idx];
        o = make_double2(s.x + p.a * (1.5 * nl.x - 0.5 * nlp.x), s.y + p.a * (1.5 * nl.y - 0.5 * nlp.y));
        break;
      }
      case 4:
        o = make_double2(p.a * y.x, p.a * y.y);
        break;
      case 5: {
        cplx x = p.x1[idx];
        o = make_double2(p.a * (y.x + p.b * (p.c * x.x)), p.a * (y.y + p.b * (p.c * x.y)));
        break;
      }
      case 6: {
        cplx sp = p.x1[idx], s2 = p.x2[idx], s = p.x3[idx];
        o = make_double2((sp.x + p.b * s2.x) + p.a * s.x, (sp.y + p.b * s2.y) + p.a * s.y);
        break;
      }
      case 9: {
        cplx nl = p.x1[idx], nlp = p.x2[idx], sv = p.x3[idx];
        const double tr = y.x + p.d * (1.5 * nl.x - 0.5 * nlp.x), ti = y.y + p.d * (1.5 * nl.y - 0.5 * nlp.y);
        o = make_double2(p.a * (tr + p.b * (p.c * sv.x)), p.a * (ti + p.b * (p.c * sv.y)));
        break;
      }
      case 8: {
        cplx x1 = p.x1[idx], x2 = p.x2[idx], x3 = p.x3[idx], x4 = p.x4[idx];
        o = make_double2(y.x + p.a * (1.5 * (x1.x + p.c * x2.x) - 0.5 * (x3.x + p.d * x4.x)),
                         y.y + p.a * (1.5 * (x1.y + p.c * x2.y) - 0.5 * (x3.y + p.d * x4.y)));
        break;
      }
      default: {
        cplx x1 = p.x1[idx], x2 = p.x2[idx];
        o = make_double2(y.x + p.a * (x1.x + p.c * x2.x), y.y + p.a * (x1.y + p.c * x2.y));
        break;
      }
    }
    p.y[idx] = o;
  }
}

int launch_lincomb(const LinArgs &p, cudaStream_t st) {
  if (!p.n) return MLEGS_OK;
  // fields read (incl. y where the mode needs it) + one written
  static const int reads[10] = {2, 2, 2, 3, 1, 2, 3, 3, 5, 4};
  prof_begin("lincomb", st, 16.0 * (double)p.n * (reads[p.mode < 0 || p.mode > 9 ? 7 : p.mode] + 1));
  lincomb_kernel<<<ew_grid(p.n), EW_THREADS, 0, st>>>(p);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// rows i < nr scaled by r(i) or 1/r(i)   (vec2tp ops:1337-1355, tp2vec ops:1509-1527)
__global__ void rscale_kernel(cplx *e, int nrl, size_t ncols, int r0, int nr, const double *r, int divide) {
  const size_t n = (size_t)nrl * ncols;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    int i = (int)(idx % nrl);
    int ig = r0 + i;
    if (ig < nr) {
      cplx v = e[idx];
      double rv = r[ig];
      e[idx] = divide ? make_double2(v.x / rv, v.y / rv) : make_double2(v.x * rv, v.y * rv);
    }
  }
}
int launch_rscale(cplx *e, int nrl, size_t ncols, int r0, int nr, const double *r, int divide, cudaStream_t st) {
  size_t n = (size_t)nrl * ncols;
  prof_begin("rscale", st, 32.0 * (double)nrl * ncols);
  rscale_kernel<<<ew_grid(n), EW_THREADS, 0, st>>>(e, nrl, ncols, r0, nr, r, divide);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// all(ieee_is_finite(e))  (check_stability, apps/vortical_flow_3d.f90:397-409)
__global__ void finite_kernel(const cplx *e, size_t n, int *flag) {
  int bad = 0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    cplx v = e[idx];
    if (!isfinite(v.x) || !isfinite(v.y)) bad = 1;
  }
  if (bad) atomicOr(flag, 1);
}
int launch_finite(const cplx *e, size_t n, int *flag, cudaStream_t st) {
  finite_kernel<<<ew_grid(n), EW_THREADS, 0, st>>>(e, n, flag);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// ---------------------------------------------------------------------------------------------
// vec2tp (ops:1413-1435): psi = -iu*mv*w1 - w2 ; chi = iu*kv*w3 + mv*kv*w4 - w5 on rows < nn of the
// retained (m,k) columns;  tp2vec (ops:1488-1502): ur = iu*mv*psi + iu*kv*ur ; up = -up - mv*kv*uz
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool k_retained(int k, int nzc, int nzcu) { return (k + 1 <= nzc) || (k + 1 >= nzcu); }

__global__ void tp_combine_kernel(TpCombineArgs a) {
  const size_t n = (size_t)a.nrl * a.npl * a.nzl;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    int i = (int)(idx % a.nrl);
    size_t t = idx / a.nrl;
    int j = (int)(t % a.npl);
    int k = (int)(t / a.npl);
    int m = a.m0 + j * a.ms;
    if (m >= a.npc) continue;
    int nn = min(a.nrc, a.nrc - m);
    if (i >= nn || !k_retained(k, a.nzc, a.nzcu)) continue;
    double mv = (double)m, kv = a.ak[k];
    cplx t1 = a.t[idx], d = a.dst[idx];
    switch (a.mode) {
      case 0:   // dst = (-iu*mv)*t
        d = make_double2(mv * t1.y, -(mv * t1.x));
        break;
      case 1:   // dst = dst - t
        d = make_double2(d.x - t1.x, d.y - t1.y);
        break;
      case 2:   // dst = (iu*kv)*t
        d = make_double2(-(kv * t1.y), kv * t1.x);
        break;
      case 3: { // dst = dst + (mv*kv)*t
        double mk = mv * kv;
        d = make_double2(d.x + mk * t1.x, d.y + mk * t1.y);
        break;
      }
      default:  // dst = dst - t
        d = make_double2(d.x - t1.x, d.y - t1.y);
        break;
    }
    a.dst[idx] = d;
  }
}
int launch_tp_combine(const TpCombineArgs &a, cudaStream_t st) {
  size_t n = (size_t)a.nrl * a.npl * a.nzl;
  prof_begin("tp_combine", st, ((a.mode == 0 || a.mode == 2) ? 32.0 : 48.0) *
                                   retained_elems(a.nrl, a.npl, a.nzl, 0, a.m0, a.nrc, a.npc, a.nzc, a.nzcu, a.ms));
  tp_combine_kernel<<<ew_grid(n), EW_THREADS, 0, st>>>(a);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// real-part updates of one column:
//   mode 0: col(i) = (col(i) - v1(i)) - v2(i)        (ihelmp log-term, ops:950)
//   mode 1: col(i) = col(i) + s*(1 + v1(i))          (vec2tp, ops:1372)
__global__ void col_update_kernel(cplx *col, int n, int mode, const double *v1, const double *v2, double s) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    cplx v = col[i];
    if (mode == 0)
      v.x = (v.x - v1[i]) - v2[i];
    else
      v.x = v.x + s * (1.0 + v1[i]);
    col[i] = v;
  }
}
int launch_col_update(cplx *col, int n, int mode, const double *v1, const double *v2, double s, cudaStream_t st) {
  if (n <= 0) return MLEGS_OK;
  col_update_kernel<<<(n + 255) / 256, 256, 0, st>>>(col, n, mode, v1, v2, s);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// ---------------------------------------------------------------------------------------------
// fftreat, ops:1002-1063: the part between the radial synthesis and the radial analysis.  Per retained (m,k) line:
// e(:nr) *= (1-x)^2; rows ns0.. zeroed for m >= 1 (on EVERY plane, ops:1031-1035); five passes of smooth()
// (ops:2149-2175) over the tail e(ns:nrdim); e(:nr) /= (1-x)^2.  One CTA per line; the tail lives in shared memory.
// smooth() accumulates into each output point in a fixed order (the last point's spread, then the centres i-1, i,
// i+1 ascending); a thread per output point adds its three contributions in that order, so results round like the
// reference's sequential loop.
// ---------------------------------------------------------------------------------------------
#define FFTREAT_MAXTAIL 1024
__global__ void __launch_bounds__(128) fftreat_tail_kernel(FftreatArgs a) {
  __shared__ cplx buf[2][FFTREAT_MAXTAIL];
  const int line = blockIdx.x;
  const int j = line % a.npl, k = line / a.npl;
  const int m = a.m0 + j * a.ms;
  if (m >= a.npc) return;
  cplx *col = a.e + ((size_t)k * a.npl + j) * a.nrl;
  const bool kept = (k < a.nzc) || (k + 1 >= a.nzcu);
  const int t0 = a.ns - 1;                 // first tail row (0-based)
  const int ni = a.nrl - t0;
  if (!kept) {
    if (m >= 1)
      for (int i = a.ns0 - 1 + threadIdx.x; i < a.nrl; i += blockDim.x) col[i] = make_double2(0.0, 0.0);
    return;
  }
  for (int i = threadIdx.x; i < a.nrl; i += blockDim.x) {
    cplx v = col[i];
    double f = 1.0;
    if (i < a.nr) {
      const double t = 1.0 - a.x[i];
      f = t * t;
      v = make_double2(v.x * f, v.y * f);
    }
    if (m >= 1 && i >= a.ns0 - 1) v = make_double2(0.0, 0.0);
    if (i >= t0)
      buf[0][i - t0] = v;
    else
      col[i] = make_double2(v.x / f, v.y / f);      // rows below the tail: (e * f) / f as the reference computes it
  }
  __syncthreads();
  int cur = 0;
  for (int pass = 0; pass < 5; ++pass) {
    const cplx *in = buf[cur];
    cplx *out = buf[cur ^ 1];
    for (int q = threadIdx.x; q < ni; q += blockDim.x) {
      cplx o = make_double2(0.0, 0.0);
      if (q == 0) o = in[0];
      if (q >= ni - 3) {
        const double w = (q == ni - 3) ? 0.2 : (q == ni - 2 ? 0.3 : 0.5);
        o = make_double2(o.x + in[ni - 1].x * w, o.y + in[ni - 1].y * w);
      }
#pragma unroll
      for (int d = -1; d <= 1; ++d) {
        const int c = q + d;               // centre (0-based) contributing to point q
        if (c < 1 || c > ni - 2) continue;
        const double f = (1.0 + (double)(ni - (c + 1)) / ((double)ni - 1.0)) * 0.5;
        const double w = (d == 0) ? f : (1.0 - f) / 2.0;
        o = make_double2(o.x + in[c].x * w, o.y + in[c].y * w);
      }
      out[q] = o;
    }
    __syncthreads();
    cur ^= 1;
  }
  for (int q = threadIdx.x; q < ni; q += blockDim.x) {
    const int i = t0 + q;
    cplx v = buf[cur][q];
    if (i < a.nr) {
      const double t = 1.0 - a.x[i];
      const double f = t * t;
      v = make_double2(v.x / f, v.y / f);
    }
    col[i] = v;
  }
}

int launch_fftreat_tail(const FftreatArgs &a, cudaStream_t st) {
  if (a.nrl - (a.ns - 1) > FFTREAT_MAXTAIL) return fail(MLEGS_E_ARG, "fftreat: radial size too large");
  if (a.nrl - (a.ns - 1) < 3) return fail(MLEGS_E_ARG, "smooth: the input length must be longer than or equal to 3.");
  const int lines = a.npl * a.nzl;
  if (lines <= 0) return MLEGS_OK;
  prof_begin("fftreat_tail", st, 32.0 * (double)a.nrl * lines);
  fftreat_tail_kernel<<<lines, 128, 0, st>>>(a);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

__global__ void tv_combine_kernel(TvCombineArgs a) {
  const size_t n = (size_t)a.nrl * a.npl * a.nzl;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    int i = (int)(idx % a.nrl);
    size_t t = idx / a.nrl;
    int j = (int)(t % a.npl);
    int k = (int)(t / a.npl);
    int m = a.m0 + j * a.ms;
    if (m >= a.npc) continue;
    int nn = min(a.nrc, a.nrc - m);
    if (i >= nn || !k_retained(k, a.nzc, a.nzcu)) continue;
    double mv = (double)m, kv = a.ak[k];
    cplx ps = a.psi[idx], ur = a.ur[idx], up = a.up[idx], uz = a.uz[idx];
    // iu*mv*psi + iu*kv*ur
    a.ur[idx] = make_double2(-(mv * ps.y) + -(kv * ur.y), mv * ps.x + kv * ur.x);
    double mk = mv * kv;
    a.up[idx] = make_double2(-up.x - mk * uz.x, -up.y - mk * uz.y);
  }
}
int launch_tv_combine(const TvCombineArgs &a, cudaStream_t st) {
  size_t n = (size_t)a.nrl * a.npl * a.nzl;
  prof_begin("tv_combine", st, 96.0 * retained_elems(a.nrl, a.npl, a.nzl, 0, a.m0, a.nrc, a.npc, a.nzc, a.nzcu, a.ms));
  tv_combine_kernel<<<ew_grid(n), EW_THREADS, 0, st>>>(a);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

}  // namespace mlegs
