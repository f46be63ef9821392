// On-device initial conditions and diagnostics of the vortical-flow apps:
//   qvort_dist_tp   /root/reference/src/apps/vortical_flow_3d.f90:258-326
//   uniform_z_fld   /root/reference/src/apps/vortical_flow_3d.f90:328-351
//   save_vort_mag   /root/reference/src/apps/vortical_flow_3d.f90:411-447 (the vorticity-magnitude field)
// The reference builds every one of them as a GLOBAL array on rank 0 and scatters it (`disassemble`), which caps the
// problem size by one host's memory; here every rank fills its own slab in HBM.  Compiled with -fmad=false so the
// arithmetic rounds like the reference's Fortran expressions; cos/sin of the collocation angles come from a host
// table (libm, like the reference), so only exp/sqrt differ from the host math library (<= 1 ulp each).
#include <cmath>
#include <cstring>
#include <vector>

#include "kernels.h"

namespace mlegs {

#define APP_THREADS 256
#define APP_MAXC 8

static inline unsigned app_grid(size_t n) {
  size_t g = (n + APP_THREADS - 1) / APP_THREADS;
  const size_t cap = 148 * 16;
  return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

struct GaussArgs {
  cplx *e;
  int nrl, npl, nzl, r0;         // local PPP block (r sharded), global index of local row 0
  int nr, nph, nz;               // physical extents: rows < nr, packed columns < np/2, planes < nz
  const double *r, *x;           // nr radial collocation points r_i and mapped x_i
  const double *cossin;          // [2*np]: cos(p_j), j < np, then sin(p_j)
  int np;
  int nc;
  double xo[APP_MAXC], yo[APP_MAXC];
  double mul, div;               // value = ((-exp(-d^2)) * mul / div) / (1 - x)^2
  double ell, noise;
  unsigned long long seed;
};

// counter-based generator for the optional perturbation (the reference's rand() is seeded from the clock,
// apps/vortical_flow_3d.f90:467-478, so its stream is not reproducible either): uniform in (-1, 1), a function of
// (seed, global element index, lane) only, i.e. independent of the decomposition
__device__ __forceinline__ double noise_pm1(unsigned long long seed, unsigned long long idx, int lane) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (2ull * idx + (unsigned long long)lane + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return ((double)(z >> 11) + 0.5) * (2.0 / 9007199254740992.0) - 1.0;
}

__global__ void gauss_vortices_kernel(GaussArgs a) {
  const size_t plane = (size_t)a.nrl * a.npl;
  // one thread per (i, j): the value does not depend on k, so it is computed once and stored to every plane
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < plane; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % a.nrl), j = (int)(idx / a.nrl);
    const int ig = a.r0 + i;
    cplx v = make_double2(0.0, 0.0);
    bool inside = false;
    if (ig < a.nr && j < a.nph) {
      const double ri_ = a.r[ig];
      const double om = 1.0 - a.x[ig];
      const double den = om * om;
      const double cr = a.cossin[2 * j], sr = a.cossin[a.np + 2 * j];           // angle p(2j-1) -> real lane
      const double ci = a.cossin[2 * j + 1], si = a.cossin[a.np + 2 * j + 1];   // angle p(2j)   -> imaginary lane
      double rr = 0.0, ri = 0.0;
      for (int c = 0; c < a.nc; ++c) {
        const double dxr = ri_ * cr - a.xo[c], dyr = ri_ * sr - a.yo[c];
        const double dxi = ri_ * ci - a.xo[c], dyi = ri_ * si - a.yo[c];
        rr = sqrt(dxr * dxr + dyr * dyr);
        ri = sqrt(dxi * dxi + dyi * dyi);
        v.x = v.x + (-exp(-(rr * rr))) * a.mul / a.div / den;
        v.y = v.y + (-exp(-(ri * ri))) * a.mul / a.div / den;
      }
      inside = (rr < a.ell) || (ri < a.ell);   // distances to the LAST centre, like the reference's loop
    }
    for (int k = 0; k < a.nzl; ++k) {
      cplx o = (k < a.nz) ? v : make_double2(0.0, 0.0);
      if (a.noise != 0.0 && inside && k < a.nz) {
        const unsigned long long g = ((unsigned long long)k * a.npl + j) * 1048576ull + (unsigned long long)ig;
        o.x = o.x + noise_pm1(a.seed, g, 0) * a.noise;
        o.y = o.y + noise_pm1(a.seed, g, 1) * a.noise;
      }
      a.e[(size_t)k * plane + idx] = o;
    }
  }
}

__global__ void fill_physical_kernel(cplx *e, int nrl, int npl, int nzl, int r0, int nr, int nph, int nz, double re,
                                     double im) {
  const size_t n = (size_t)nrl * npl * nzl;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % nrl);
    const size_t t = idx / nrl;
    const int j = (int)(t % npl), k = (int)(t / npl);
    const bool in = (r0 + i < nr) && (j < nph) && (k < nz);
    e[idx] = in ? make_double2(re, im) : make_double2(0.0, 0.0);
  }
}

// vormag%e = cmplx(sqrt(Re(wr)^2 + Re(wp)^2 + Re(wz)^2), sqrt(Im(..)^2 ...)), every local element (:435-447)
__global__ void vecmag_kernel(cplx *out, const cplx *wr, const cplx *wp, const cplx *wz, size_t n) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const cplx a = wr[idx], b = wp[idx], c = wz[idx];
    out[idx] = make_double2(sqrt(a.x * a.x + b.x * b.x + c.x * c.x), sqrt(a.y * a.y + b.y * b.y + c.y * c.y));
  }
}

static cudaStream_t strm() { return (cudaStream_t)ctx().stream; }
static bool is_space(const mlegs_field *f, const char *sp) { return strncmp(f->space, sp, 3) == 0; }

static int cossin_table() {
  Context &c = ctx();
  if (c.d_cossin_p) return MLEGS_OK;
  const int np = c.p.np;
  std::vector<double> h(2 * (size_t)np);
  const double pi = std::acos(-1.0);
  for (int j = 0; j < np; ++j) {
    const double ang = 2 * pi / np * j;          // tfm%p, sinit:95
    h[j] = std::cos(ang);
    h[np + j] = std::sin(ang);
  }
  CUDA_TRY(cudaMalloc((void **)&c.d_cossin_p, h.size() * sizeof(double)));
  CUDA_TRY(cudaMemcpy(c.d_cossin_p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  return MLEGS_OK;
}

int trans_impl(mlegs_field *s, const char *to);
int delsqp_impl(mlegs_field *s, bool inverse);
int zeroat1_impl(mlegs_field *s);
int tp2curlvec_impl(const mlegs_field *psi, const mlegs_field *chi, mlegs_field *wr, mlegs_field *wp, mlegs_field *wz);

static int gauss_impl(mlegs_field *s, int nc, const double *xo, const double *yo, double mul, double div, double noise,
                      unsigned long long seed) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  if (!is_space(s, "PPP")) return fail(MLEGS_E_ARG, "gauss_vortices: scalar must be in PPP");
  if (nc < 0 || nc > APP_MAXC) return fail(MLEGS_E_ARG, "gauss_vortices: at most 8 centres");
  if (div == 0.0) return fail(MLEGS_E_ARG, "gauss_vortices: zero divisor");
  MLEGS_TRY(cossin_table());
  GaussArgs a;
  a.e = (cplx *)s->e;
  a.nrl = s->loc_sz[0];
  a.npl = s->loc_sz[1];
  a.nzl = s->loc_sz[2];
  a.r0 = s->loc_st[0];
  a.nr = c.p.nr;
  a.nph = c.p.np / 2;
  a.nz = c.p.nz;
  a.r = c.d_r;
  a.x = c.d_x;
  a.cossin = c.d_cossin_p;
  a.np = c.p.np;
  a.nc = nc;
  for (int i = 0; i < nc; ++i) {
    a.xo[i] = xo[i];
    a.yo[i] = yo[i];
  }
  a.mul = mul;
  a.div = div;
  a.ell = c.p.ell;
  a.noise = noise;
  a.seed = seed;
  const size_t plane = (size_t)a.nrl * a.npl;
  if (plane == 0 || a.nzl == 0) return MLEGS_OK;
  prof_begin("gauss_vortices", strm());
  gauss_vortices_kernel<<<app_grid(plane), APP_THREADS, 0, strm()>>>(a);
  prof_end(strm());
  KERNEL_CHECK();
  return MLEGS_OK;
}

}  // namespace mlegs

using namespace mlegs;

extern "C" {

int mlegs_b200_gauss_vortices(mlegs_field *s, int ncentres, const double *xo, const double *yo, double mul, double div,
                              double ran_noise, unsigned long long seed) {
  return gauss_impl(s, ncentres, xo, yo, mul, div, ran_noise, seed);
}

int mlegs_b200_fill_physical(mlegs_field *s, double re, double im) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  if (!is_space(s, "PPP")) return fail(MLEGS_E_ARG, "uniform_z_fld: scalar must be in PPP");
  const size_t n = (size_t)s->loc_sz[0] * s->loc_sz[1] * s->loc_sz[2];
  if (n == 0) return MLEGS_OK;
  prof_begin("fill_physical", strm());
  fill_physical_kernel<<<app_grid(n), APP_THREADS, 0, strm()>>>((cplx *)s->e, s->loc_sz[0], s->loc_sz[1], s->loc_sz[2],
                                                               s->loc_st[0], c.p.nr, c.p.np / 2, c.p.nz, re, im);
  prof_end(strm());
  KERNEL_CHECK();
  return MLEGS_OK;
}

int mlegs_b200_qvort_dist_tp(mlegs_field *psi, mlegs_field *chi, double q, double ran_noise, unsigned long long seed) {
  if (!is_space(psi, "FFF") || !is_space(chi, "FFF"))
    return fail(MLEGS_E_ARG, "qvort_dist_tp: psi and chi must be in FFF");
  if (q == 0.0) return fail(MLEGS_E_ARG, "qvort_dist_tp: q must not be zero");
  const double xo[2] = {-2.0, 2.0}, yo[2] = {0.0, 0.0};   // do xo = -2, 2, 4; yo = 0
  // the reference transforms the (zero) inputs to PPP only to obtain the physical layout (:275-276)
  field_set_layout(psi, true);
  field_set_layout(chi, true);
  memcpy(psi->space, "PPP", 4);
  memcpy(chi->space, "PPP", 4);
  MLEGS_TRY(gauss_impl(psi, 2, xo, yo, 2.0, 1.0, ran_noise, seed));
  MLEGS_TRY(trans_impl(psi, "FFF"));
  MLEGS_TRY(delsqp_impl(psi, true));
  MLEGS_TRY(zeroat1_impl(psi));
  MLEGS_TRY(gauss_impl(chi, 2, xo, yo, 1.0, q, ran_noise, seed + 0x51ED270B7F4A7C15ull));
  MLEGS_TRY(trans_impl(chi, "FFF"));
  MLEGS_TRY(delsqp_impl(chi, true));
  MLEGS_TRY(zeroat1_impl(chi));
  return MLEGS_OK;
}

int mlegs_b200_vort_mag(const mlegs_field *psi, const mlegs_field *chi, mlegs_field *wr, mlegs_field *wp,
                        mlegs_field *wz, mlegs_field *vormag) {
  MLEGS_TRY(tp2curlvec_impl(psi, chi, wr, wp, wz));
  if (!is_space(vormag, "PPP")) return fail(MLEGS_E_ARG, "save_vort_mag: vormag must be in PPP");
  const size_t n = (size_t)wr->loc_sz[0] * wr->loc_sz[1] * wr->loc_sz[2];
  if (n == 0) return MLEGS_OK;
  prof_begin("vecmag", strm());
  vecmag_kernel<<<app_grid(n), APP_THREADS, 0, strm()>>>((cplx *)vormag->e, (const cplx *)wr->e, (const cplx *)wp->e,
                                                        (const cplx *)wz->e, n);
  prof_end(strm());
  KERNEL_CHECK();
  return MLEGS_OK;
}

}  // extern "C"
