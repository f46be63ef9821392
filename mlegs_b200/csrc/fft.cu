// Batched strided line FFTs for the azimuthal (real, packed two-per-complex) and axial
// (complex) directions.  Replaces the per-line FFTE calls of
// /root/reference/src/submodules/mlegs_scalar_ops.f90:1567-1848 (dzfft2d / zdfft2d / zfft1d).
//
// Layout: the field is e(i, j, k) column-major; a phi-line runs along j (stride = rows), a z-line
// along k (stride = rows*cols).  In both cases the radial/row index i is the fastest index and is
// the batch direction, so a CTA stages `ti` neighbouring lines -- ti consecutive complex(16 B)
// per point -- through shared memory with fully coalesced 16-byte accesses, runs a mixed-radix
// (2,3,4,5) Stockham autosort FFT on all of them at once in shared memory (line index fastest, so
// shared accesses are conflict free), and writes the tile back.  HBM-bound: every element is
// read once and written once.
#include <cmath>
#include <cstdio>

#include "dist_dev.cuh"

namespace mlegs {

#define FFT_THREADS 256
#define FFT_MAXPASS 16

struct FftPassList {
  int npass;
  int radix[FFT_MAXPASS];
};

__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
// multiply by -i
__device__ __forceinline__ cplx mul_mi(cplx a) { return make_double2(a.y, -a.x); }

template <int R>
__device__ __forceinline__ void dft(cplx *v);

template <>
__device__ __forceinline__ void dft<2>(cplx *v) {
  cplx a = v[0], b = v[1];
  v[0] = cadd(a, b);
  v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void dft<3>(cplx *v) {
  const double s3 = 0.86602540378443864676;  // sqrt(3)/2
  cplx t1 = cadd(v[1], v[2]);
  cplx t2 = make_double2(v[0].x - 0.5 * t1.x, v[0].y - 0.5 * t1.y);
  cplx d = csub(v[1], v[2]);
  cplx t3 = mul_mi(make_double2(s3 * d.x, s3 * d.y));
  v[0] = cadd(v[0], t1);
  v[1] = cadd(t2, t3);
  v[2] = csub(t2, t3);
}
template <>
__device__ __forceinline__ void dft<4>(cplx *v) {
  cplx a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
  cplx c = cadd(v[1], v[3]), d = mul_mi(csub(v[1], v[3]));
  v[0] = cadd(a, c);
  v[1] = cadd(b, d);
  v[2] = csub(a, c);
  v[3] = csub(b, d);
}
template <>
__device__ __forceinline__ void dft<5>(cplx *v) {
  const double c1 = 0.30901699437494742410;   // cos(2pi/5)
  const double c2 = -0.80901699437494742410;  // cos(4pi/5)
  const double s1 = 0.95105651629515357212;   // sin(2pi/5)
  const double s2 = 0.58778525229247312917;   // sin(4pi/5)
  cplx t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
  cplx t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
  cplx a1 = make_double2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
  cplx a2 = make_double2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
  cplx b1 = mul_mi(make_double2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));  // -i b1
  cplx b2 = mul_mi(make_double2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));  // -i b2
  v[0] = make_double2(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
  v[1] = cadd(a1, b1);
  v[4] = csub(a1, b1);
  v[2] = cadd(a2, b2);
  v[3] = csub(a2, b2);
}

// One Stockham pass of radix R over `ti` lines held as src[point*ti + line].
template <int R>
__device__ __forceinline__ void stockham_pass(const cplx *__restrict__ src, cplx *__restrict__ dst, int n, int ti,
                                              int ns, const cplx *__restrict__ tw, int tw_step) {
  const int nb = n / R;
  const int total = nb * ti;
  const int tw_mul = (n / (ns * R)) * tw_step;
  for (int idx = threadIdx.x; idx < total; idx += FFT_THREADS) {
    int b = idx / ti;
    int l = idx - b * ti;
    int k = b % ns;
    cplx v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) v[t] = src[(b + t * nb) * ti + l];
    if (ns > 1) {
#pragma unroll
      for (int t = 1; t < R; ++t) v[t] = cmul(v[t], __ldg(&tw[t * k * tw_mul]));
    }
    dft<R>(v);
    int j0 = (b - k) * R + k;
#pragma unroll
    for (int t = 0; t < R; ++t) dst[(j0 + t * ns) * ti + l] = v[t];
  }
}

// Forward (e^{-i}) FFT of length n on all lines of the tile.  Returns the buffer holding the result.
__device__ __forceinline__ cplx *fft_tile(cplx *a, cplx *b, int n, int ti, const FftPassList &pl,
                                          const cplx *__restrict__ tw, int tw_step) {
  int ns = 1;
  cplx *src = a, *dst = b;
  for (int p = 0; p < pl.npass; ++p) {
    int r = pl.radix[p];
    if (r == 4)
      stockham_pass<4>(src, dst, n, ti, ns, tw, tw_step);
    else if (r == 2)
      stockham_pass<2>(src, dst, n, ti, ns, tw, tw_step);
    else if (r == 3)
      stockham_pass<3>(src, dst, n, ti, ns, tw, tw_step);
    else
      stockham_pass<5>(src, dst, n, ti, ns, tw, tw_step);
    ns *= r;
    __syncthreads();
    cplx *t = src;
    src = dst;
    dst = t;
  }
  return src;
}

// MODE: FFT_C2C_FWD / FFT_C2C_BWD / FFT_R2C_FWD / FFT_C2R_BWD.
// n = complex length of the in-smem FFT (np/2 for the real modes, nz for c2c).
// tw = table of exp(-2 pi i j / tw_order), j < tw_order.
template <int MODE>
__global__ void __launch_bounds__(FFT_THREADS)
fft_lines_kernel(const cplx *__restrict__ in, cplx *__restrict__ out, int n, int ti, long long batch0,
                 long long stride_pt, long long stride_b1, const cplx *__restrict__ tw, int tw_order, double scale,
                 FftPassList pl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx *bufa = reinterpret_cast<cplx *>(smem_raw);
  cplx *bufb = bufa + (size_t)(n + 1) * ti;

  const long long b0 = (long long)blockIdx.x * ti;
  const int nl = (int)min((long long)ti, batch0 - b0);
  const cplx *gin = in + (long long)blockIdx.y * stride_b1 + b0;
  cplx *gout = out + (long long)blockIdx.y * stride_b1 + b0;
  const int tw_step = tw_order / n;   // 1 for c2c, 2 for the real modes (tw_order == 2n)
  const int npts_in = (MODE == FFT_C2R_BWD) ? n + 1 : n;

  // ---- load tile (coalesced: line index fastest) ----
  for (int idx = threadIdx.x; idx < npts_in * ti; idx += FFT_THREADS) {
    int j = idx / ti;
    int l = idx - j * ti;
    cplx v = make_double2(0.0, 0.0);
    if (l < nl) v = gin[(long long)j * stride_pt + l];
    if (MODE == FFT_C2C_BWD) v = cconj(v);
    if (MODE == FFT_C2R_BWD)
      bufb[idx] = v;
    else
      bufa[idx] = v;
  }
  __syncthreads();

  if (MODE == FFT_C2R_BWD) {
    // Hermitian half spectrum C_0..C_n  ->  packed Z_m, m < n (conjugated for the conj-FFT-conj inverse).
    // Im(C_0), Im(C_n) are ignored like external/ffte-7.0/zdfft2d.f:119-128 does.
    for (int idx = threadIdx.x; idx < n * ti; idx += FFT_THREADS) {
      int m = idx / ti;
      int l = idx - m * ti;
      cplx cm = bufb[m * ti + l];
      cplx cc = cconj(bufb[(n - m) * ti + l]);
      if (m == 0) {
        cm.y = 0.0;
        cc.y = 0.0;
      }
      cplx s = cadd(cm, cc), d = csub(cm, cc);
      cplx wm = cconj(__ldg(&tw[m]));             // e^{+2 pi i m / (2n)}
      cplx t = cmul(wm, d);                        // w^{-m} (Cm - conj C_{n-m})
      cplx z = make_double2(s.x - t.y, s.y + t.x); // s + i t
      bufa[idx] = cconj(z);
    }
    __syncthreads();
  }

  cplx *res = fft_tile(bufa, bufb, n, ti, pl, tw, tw_step);
  cplx *oth = (res == bufa) ? bufb : bufa;

  if (MODE == FFT_R2C_FWD) {
    // packed Z -> X_m, m = 0..n  (external/ffte-7.0/dzfft2d.f NY=1 branch == rfft), times scale (= 1/np)
    for (int idx = threadIdx.x; idx < (n + 1) * ti; idx += FFT_THREADS) {
      int m = idx / ti;
      int l = idx - m * ti;
      int m1 = (m == n) ? 0 : m;
      int m2 = (m == 0) ? 0 : n - m;
      cplx zm = res[m1 * ti + l];
      cplx zc = cconj(res[m2 * ti + l]);
      cplx e = make_double2(0.5 * (zm.x + zc.x), 0.5 * (zm.y + zc.y));
      cplx d = csub(zm, zc);
      cplx o = make_double2(0.5 * d.y, -0.5 * d.x);   // (-i/2) d
      cplx x = cadd(e, cmul(__ldg(&tw[m]), o));
      if (l < nl) gout[(long long)m * stride_pt + l] = make_double2(x.x * scale, x.y * scale);
    }
    return;
  }

  // ---- store tile ----
  for (int idx = threadIdx.x; idx < n * ti; idx += FFT_THREADS) {
    int j = idx / ti;
    int l = idx - j * ti;
    cplx v = res[idx];
    if (MODE == FFT_C2C_BWD || MODE == FFT_C2R_BWD) v = cconj(v);
    if (l < nl) gout[(long long)j * stride_pt + l] = make_double2(v.x * scale, v.y * scale);
  }
  if (MODE == FFT_C2R_BWD) {
    // padding column keeps the Nyquist input times np (quirk Q3; ops:1702-1705)
    const double fac = (double)(2 * n);
    for (int l = threadIdx.x; l < nl; l += FFT_THREADS) {
      // the untouched copy of C_n: bufb was consumed by the FFT ping-pong, so re-read it from global
      cplx cn = gin[(long long)n * stride_pt + l];
      (void)oth;
      gout[(long long)n * stride_pt + l] = make_double2(cn.x * fac, cn.y * fac);
    }
  }
}

int make_fft_plan(int n, int extra_points, FftPlan *plan) {
  plan->n = n;
  plan->npass = 0;
  int rem = n;
  while (rem % 4 == 0) { plan->radix[plan->npass++] = 4; rem /= 4; }
  while (rem % 2 == 0) { plan->radix[plan->npass++] = 2; rem /= 2; }
  while (rem % 3 == 0) { plan->radix[plan->npass++] = 3; rem /= 3; }
  while (rem % 5 == 0) { plan->radix[plan->npass++] = 5; rem /= 5; }
  if (rem != 1 || plan->npass > FFT_MAXPASS)
    return fail(MLEGS_E_ARG, "tfm_kit_init: np must only have factors of 2, 3 and 5");
  (void)extra_points;
  // lines per CTA: keep both ping-pong buffers within ~72 KB so two CTAs fit on an SM
  int ti = 32;
  while (ti > 4 && (size_t)2 * (n + 1) * ti * sizeof(cplx) > 72 * 1024) ti >>= 1;
  plan->ti = ti;
  plan->smem = (size_t)2 * (n + 1) * ti * sizeof(cplx);
  if (plan->smem > 220 * 1024) return fail(MLEGS_E_ARG, "fft: transform length too large for shared memory");
  return MLEGS_OK;
}

int setup_fft_kernels() {
  const int maxsm = 220 * 1024;
  CUDA_TRY(cudaFuncSetAttribute(fft_lines_kernel<FFT_C2C_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
  CUDA_TRY(cudaFuncSetAttribute(fft_lines_kernel<FFT_C2C_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
  CUDA_TRY(cudaFuncSetAttribute(fft_lines_kernel<FFT_R2C_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
  CUDA_TRY(cudaFuncSetAttribute(fft_lines_kernel<FFT_C2R_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
  return MLEGS_OK;
}

int launch_fft_lines(FftMode mode, const FftPlan &plan, const cplx *in, cplx *out, long long batch0,
                     long long stride_pt, int batch1, long long stride_b1, const double *tw, int tw_order,
                     double scale, cudaStream_t st, const FieldBatch *fb, const RowScale *rs) {
  if (batch0 <= 0 || batch1 <= 0) return MLEGS_OK;
  if (rs && rs->mode != 0 && !fft_reg_supported(plan.n))
    return fail(MLEGS_E_STATE, "fft: fused row scaling needs the register kernels");
  static const char *names[4] = {"fft_z_forward", "fft_z_backward", "fft_phi_forward", "fft_phi_backward"};
  if (fb && fb->n > 0 && !fft_reg_supported(plan.n)) {
    // lengths outside the register kernels: one launch per scalar
    for (int i = 0; i < fb->n; ++i)
      MLEGS_TRY(launch_fft_lines(mode, plan, fb->in[i], fb->out[i], batch0, stride_pt, batch1, stride_b1, tw, tw_order,
                                 scale, st, nullptr));
    return MLEGS_OK;
  }
  // algorithmic bytes: a c2c line moves 2 x 16 N bytes; an r2c / c2r line 16 N (the packed reals) + 16 (N + 1)
  const double nfld = (fb && fb->n > 0) ? fb->n : 1;
  const bool is_phi = mode == FFT_R2C_FWD || mode == FFT_C2R_BWD;
  const double line_bytes = is_phi ? 16.0 * plan.n + 16.0 * (plan.n + 1) : 32.0 * plan.n;
  const double alg_bytes = nfld * (double)batch0 * batch1 * line_bytes;
  if (fft_reg_supported(plan.n)) {
    prof_begin(names[(int)mode], st, alg_bytes);
    int rc = launch_fft_reg(mode, plan.n, in, out, batch0 * batch1, batch0, stride_b1, stride_pt, tw, tw_order, scale,
                            nullptr, 0, 0, st, nullptr, 0, fb, rs);
    prof_end(st);
    MLEGS_TRY(rc);
    KERNEL_CHECK();
    return MLEGS_OK;
  }
  FftPassList pl;
  pl.npass = plan.npass;
  for (int i = 0; i < FFT_MAXPASS; ++i) pl.radix[i] = plan.radix[i];
  dim3 grid((unsigned)((batch0 + plan.ti - 1) / plan.ti), (unsigned)batch1);
  const cplx *twc = reinterpret_cast<const cplx *>(tw);
  prof_begin(names[(int)mode], st, alg_bytes);
  switch (mode) {
    case FFT_C2C_FWD:
      fft_lines_kernel<FFT_C2C_FWD><<<grid, FFT_THREADS, plan.smem, st>>>(in, out, plan.n, plan.ti, batch0, stride_pt,
                                                                          stride_b1, twc, tw_order, scale, pl);
      break;
    case FFT_C2C_BWD:
      fft_lines_kernel<FFT_C2C_BWD><<<grid, FFT_THREADS, plan.smem, st>>>(in, out, plan.n, plan.ti, batch0, stride_pt,
                                                                          stride_b1, twc, tw_order, scale, pl);
      break;
    case FFT_R2C_FWD:
      fft_lines_kernel<FFT_R2C_FWD><<<grid, FFT_THREADS, plan.smem, st>>>(in, out, plan.n, plan.ti, batch0, stride_pt,
                                                                          stride_b1, twc, tw_order, scale, pl);
      break;
    case FFT_C2R_BWD:
      fft_lines_kernel<FFT_C2R_BWD><<<grid, FFT_THREADS, plan.smem, st>>>(in, out, plan.n, plan.ti, batch0, stride_pt,
                                                                          stride_b1, twc, tw_order, scale, pl);
      break;
  }
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

int launch_fft_phi_forward_put(const FftPlan &plan, const cplx *in, long long rows, int nz, long long plane,
                               const double *tw, int tw_order, double scale, const PeerTable &peer, int nrdim,
                               cudaStream_t st, const FieldBatch *fb, const RowScale *rs) {
  if (!fft_reg_supported(plan.n)) return fail(MLEGS_E_STATE, "fft: the fused exchange needs the register kernels");
  prof_begin("fft_phi_forward_put", st,
             ((fb && fb->n > 0) ? fb->n : 1) * (double)rows * nz * (16.0 * plan.n + 16.0 * (plan.n + 1)));
  int rc = launch_fft_reg(FFT_R2C_FWD, plan.n, in, nullptr, rows * nz, rows, plane, rows, tw, tw_order, scale, nullptr, 0, 0,
                          st, &peer, nrdim, fb, rs);
  prof_end(st);
  MLEGS_TRY(rc);
  KERNEL_CHECK();
  return MLEGS_OK;
}

int launch_fft_z_compact(FftMode mode, const FftPlan &plan, const cplx *in, cplx *out, const int *colstart, int ncols,
                         int nrl, long long nlines, long long stride_pt, const double *tw, int tw_order, double scale,
                         cudaStream_t st, const FieldBatch *fb) {
  if (nlines <= 0) return MLEGS_OK;
  if (!fft_reg_supported(plan.n)) return fail(MLEGS_E_STATE, "fft: compact mode needs the register kernels");
  prof_begin(mode == FFT_C2C_FWD ? "fft_z_forward" : "fft_z_backward", st,
             ((fb && fb->n > 0) ? fb->n : 1) * (double)nlines * 32.0 * plan.n);
  int rc = launch_fft_reg(mode, plan.n, in, out, nlines, 1, 0, stride_pt, tw, tw_order, scale, colstart, ncols, nrl, st,
                          nullptr, 0, fb);
  prof_end(st);
  MLEGS_TRY(rc);
  KERNEL_CHECK();
  return MLEGS_OK;
}

}  // namespace mlegs
