// Register-resident batched line FFTs for power-of-two lengths (32 .. 1024 complex points): the fast
// path behind launch_fft_lines for the azimuthal (real, packed) and axial (complex) transforms of
// /root/reference/src/submodules/mlegs_scalar_ops.f90:1567-1848 (dzfft2d / zdfft2d / zfft1d per line).
//
// A line of N points is owned by T = N/E threads, each holding E points in registers.  The transform
// is a Stockham autosort FFT in 2 (N <= 256) or 3 passes of radix <= 16; passes exchange data through
// one shared-memory tile laid out [point][line] so that every shared access of a quarter warp is one
// contiguous 128-byte wavefront (conflict free).  The first pass loads straight from HBM and the last
// pass stores straight to HBM: lanes map to neighbouring lines, which are neighbouring 16-byte elements
// in memory (the radial index is the fastest one), so global accesses are coalesced 16-byte vectors and
// every element is read once and written once.  All E loads of a thread are issued before the first
// butterfly (memory-level parallelism = E x 16 B per thread).
#include <algorithm>
#include <cstdlib>

#include "dist_dev.cuh"

namespace mlegs {

namespace {

__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cplx mul_mi(cplx a) { return make_double2(a.y, -a.x); }   // * (-i)
// * exp(-i pi/4) = (1 - i)/sqrt(2);  * exp(-3 i pi/4) = (-1 - i)/sqrt(2)
#define RSQRT2 0.70710678118654752440
__device__ __forceinline__ cplx mul_w8_1(cplx a) { return make_double2((a.x + a.y) * RSQRT2, (a.y - a.x) * RSQRT2); }
__device__ __forceinline__ cplx mul_w8_3(cplx a) { return make_double2((a.y - a.x) * RSQRT2, -(a.x + a.y) * RSQRT2); }
#define C16_1 0.92387953251128675613   // cos(pi/8)
#define S16_1 0.38268343236508977173   // sin(pi/8)

// natural-order forward (e^{-i}) DFTs of a register array
template <int R>
__device__ __forceinline__ void dft(cplx *v);

template <>
__device__ __forceinline__ void dft<2>(cplx *v) {
  cplx a = v[0], b = v[1];
  v[0] = cadd(a, b);
  v[1] = csub(a, b);
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
__device__ __forceinline__ void dft<8>(cplx *v) {
  // decimation in frequency: a_i = v_i + v_{i+4}, b_i = (v_i - v_{i+4}) W8^i; X_{2j} = DFT4(a)_j, X_{2j+1} = DFT4(b)_j
  cplx a[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    a[i] = cadd(v[i], v[i + 4]);
    b[i] = csub(v[i], v[i + 4]);
  }
  b[1] = mul_w8_1(b[1]);
  b[2] = mul_mi(b[2]);
  b[3] = mul_w8_3(b[3]);
  dft<4>(a);
  dft<4>(b);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = a[j];
    v[2 * j + 1] = b[j];
  }
}
template <>
__device__ __forceinline__ void dft<16>(cplx *v) {
  cplx a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = cadd(v[i], v[i + 8]);
    b[i] = csub(v[i], v[i + 8]);
  }
  // W16^i = exp(-i pi i/8)
  b[1] = cmul(b[1], make_double2(C16_1, -S16_1));
  b[2] = mul_w8_1(b[2]);
  b[3] = cmul(b[3], make_double2(S16_1, -C16_1));
  b[4] = mul_mi(b[4]);
  b[5] = cmul(b[5], make_double2(-S16_1, -C16_1));
  b[6] = mul_w8_3(b[6]);
  b[7] = cmul(b[7], make_double2(-C16_1, -S16_1));
  dft<8>(a);
  dft<8>(b);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[2 * j] = a[j];
    v[2 * j + 1] = b[j];
  }
}

// One Stockham pass of radix R on the register tile.  Before: v[j] = x[t + T j].  After: v[q + u Q] holds
// y[out_index(q, u)], out_index = (b - k) R + k + u NS with b = t + q T, k = b mod NS.
template <int N, int E, int R, int NS>
__device__ __forceinline__ void pass_compute(cplx (&v)[E], int t, const cplx *__restrict__ tw, int tw_unit) {
  constexpr int T = N / E, Q = E / R;
  static_assert(E % R == 0 && N % (NS * R) == 0, "bad pass");
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    cplx x[R];
#pragma unroll
    for (int u = 0; u < R; ++u) x[u] = v[q + u * Q];
    if (NS > 1) {
      const int k = (t + q * T) & (NS - 1);
      const int step = k * (N / (NS * R)) * tw_unit;
#pragma unroll
      for (int u = 1; u < R; ++u) x[u] = cmul(x[u], __ldg(&tw[u * step]));
    }
    dft<R>(x);
#pragma unroll
    for (int u = 0; u < R; ++u) v[q + u * Q] = x[u];
  }
}

template <int N, int E, int R, int NS>
__device__ __forceinline__ int out_index(int t, int q, int u) {
  constexpr int T = N / E;
  const int b = t + q * T;
  const int k = b & (NS - 1);
  return (b - k) * R + k + u * NS;
}

template <int N, int E, int R, int NS, int L>
__device__ __forceinline__ void pass_to_smem(const cplx (&v)[E], cplx *sm, int t, int l) {
  constexpr int Q = E / R;
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int u = 0; u < R; ++u) sm[out_index<N, E, R, NS>(t, q, u) * L + l] = v[q + u * Q];
}

template <int N, int E, int L>
__device__ __forceinline__ void smem_to_regs(cplx (&v)[E], const cplx *sm, int t, int l) {
  constexpr int T = N / E;
#pragma unroll
  for (int j = 0; j < E; ++j) v[j] = sm[(t + T * j) * L + l];
}

// radix schedule: R0 = E; R1 = min(E, N/E); R2 = N/(R0 R1) (1 when two passes suffice)
template <int N, int E>
struct Sched {
  static constexpr int R0 = E;
  static constexpr int R1 = (N / E) < E ? (N / E) : E;
  static constexpr int R2 = N / (R0 * R1);
  static_assert(R2 == 1 || R2 == 2 || R2 == 4 || R2 == 8 || R2 == 16, "unsupported length");
};

}  // namespace

struct FftRegArgs {
  const cplx *in[MLEGS_MAXB];   // one entry per scalar of the launch (blockIdx.y)
  cplx *out[MLEGS_MAXB];
  long long nlines;      // number of lines handled by this launch
  long long batch0;      // lines q .. with q / batch0 equal are contiguous in memory
  long long stride_b1;   // element offset between such runs
  long long stride_pt;   // element stride between consecutive points of a line
  const cplx *tw;        // exp(-2 pi i j / tw_order)
  int tw_order;
  double scale;
  const int *colstart;   // compact mode (axial FFT of the retained lines only): prefix sums of nn(m) per column
  int ncols, nrl;
  // fused exchange(2,1) (FFT_R2C_FWD on several ranks): output column m goes straight into the window of the rank
  // that owns m, at row r_off[me] + i of its (nrdim, m_cnt, nz) block; the kernel ends with the exchange barrier
  int use_peer;
  int nrdim;
  RowScale rs;           // fused r*u (r2c loads) / u/r (c2r stores) of the scalars in rs.mask
  PeerTable pt;
  int dbg_copy;          // diagnostic (MLEGS_FFT_COPY=1): move the data with the kernel's access pattern, no transform
};

// x / b with the reciprocal y = RN(1/b) computed once per line: q = RN(x y), r = x - b q (exact in an fma),
// q' = RN(q + r y) is the correctly rounded quotient (Markstein), i.e. what `x / b` returns, at 3 instead of ~10 FP64
// instructions per value.
__device__ __forceinline__ double div_by(double x, double b, double y) {
  const double q = x * y;
  const double r = fma(-b, q, x);
  return fma(r, y, q);
}

// MODE: FFT_C2C_FWD / FFT_C2C_BWD / FFT_R2C_FWD / FFT_C2R_BWD (kernels.h).
// one tile = L lines of one scalar (blockIdx.y); `blk` = tile index
// EXT != 0: the launch carries a fused exchange and/or fused row scaling (kept out of the plain kernels: the extra state
// costs them 8-26 registers and a resident CTA per SM); EXT == 2: c2r whose input columns are in the transit layout of a
// fused exchange(1,2) (a kernel of its own: the column arithmetic costs the row-scaling kernel 15 %)
template <int MODE, int N, int E, int THREADS, int EXT>
__device__ __forceinline__ void fft_reg_tile(const FftRegArgs &a, long long blk) {
  constexpr int T = N / E, L = THREADS / T;
  using S = Sched<N, E>;
  // declared here, not passed in: a pointer parameter loses the shared address space (generic loads, 64-bit addresses,
  // +16 registers on the r2c kernel)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx *sm = reinterpret_cast<cplx *>(smem_raw);
  const int tid = threadIdx.x;
  const int l = tid % L, t = tid / L;
  const long long q = blk * L + l;
  const bool ok = q < a.nlines;
  long long base = 0;
  if (ok) {
    if (a.colstart) {
      // column j with colstart[j] <= q < colstart[j+1]
      int lo = 0, hi = a.ncols;
      while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if ((long long)__ldg(&a.colstart[mid]) <= q) lo = mid; else hi = mid;
      }
      base = (long long)lo * a.nrl + (q - __ldg(&a.colstart[lo]));
    } else {
      const long long run = q / a.batch0;
      base = run * a.stride_b1 + (q - run * a.batch0);
    }
  }
  const cplx *gin = a.in[blockIdx.y] + base;
  cplx *gout = a.out[blockIdx.y] + base;
  // fused row scaling (real modes only: lines are (row, plane) pairs, row = q mod batch0)
  double rv = 1.0, rinv = 1.0;
  bool scaled = false;
  if (EXT && (MODE == FFT_R2C_FWD || MODE == FFT_C2R_BWD) && a.rs.mode != 0 && ((a.rs.mask >> blockIdx.y) & 1u) && ok) {
    const int ig = a.rs.r0 + (int)(q % a.batch0);
    if (ig < a.rs.nr) {
      rv = __ldg(&a.rs.r[ig]);
      if (MODE == FFT_C2R_BWD) rinv = 1.0 / rv;
      scaled = true;
    }
  }
  const int tw_unit = a.tw_order / N;   // 1 for c2c; 2 for the real modes (tw_order == 2N)
  const cplx zero = make_double2(0.0, 0.0);

  cplx v[E];
  cplx nyq = zero;   // C_N of the Hermitian input (c2r only)

  if (MODE == FFT_C2R_BWD) {
    // Hermitian half spectrum C_0..C_N -> shared, then the packed Z_m (conjugated for the conj-FFT-conj inverse).
    // Im(C_0), Im(C_N) are ignored like external/ffte-7.0/zdfft2d.f:119-128 does.
    // column of point m inside the input block (transit layout of a fused exchange: grouped by owner of m)
    auto colof = [&](int m) -> long long {
      if (EXT == 2) return (long long)a.rs.perm_off[m % a.rs.perm_p] + m / a.rs.perm_p;
      return m;
    };
#pragma unroll
    for (int j = 0; j < E; ++j) v[j] = ok ? gin[colof(t + T * j) * a.stride_pt] : zero;
    if (t == 0) nyq = ok ? gin[colof(N) * a.stride_pt] : zero;
#pragma unroll
    for (int j = 0; j < E; ++j) sm[(t + T * j) * L + l] = v[j];
    if (t == 0) sm[N * L + l] = nyq;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const int m = t + T * j;
      cplx cm = v[j];
      cplx cc = cconj(sm[(N - m) * L + l]);
      if (m == 0) {
        cm.y = 0.0;
        cc.y = 0.0;
      }
      cplx s = cadd(cm, cc), d = csub(cm, cc);
      cplx wm = cconj(__ldg(&a.tw[m]));               // e^{+2 pi i m / (2N)}
      cplx tt = cmul(wm, d);
      v[j] = make_double2(s.x - tt.y, -(s.y + tt.x));  // conj(s + i tt)
    }
    __syncthreads();
  } else {
#pragma unroll
    for (int j = 0; j < E; ++j) {
      cplx x = ok ? gin[(long long)(t + T * j) * a.stride_pt] : zero;
      if (EXT && MODE == FFT_R2C_FWD && scaled) x = make_double2(x.x * rv, x.y * rv);
      v[j] = (MODE == FFT_C2C_BWD) ? cconj(x) : x;
    }
  }
  if (a.dbg_copy) {
#pragma unroll
    for (int j = 0; j < E; ++j)
      if (ok) gout[(long long)(t + T * j) * a.stride_pt] = v[j];
    return;
  }

  // ---- pass 0 ----
  pass_compute<N, E, S::R0, 1>(v, t, a.tw, tw_unit);
  pass_to_smem<N, E, S::R0, 1, L>(v, sm, t, l);
  __syncthreads();
  smem_to_regs<N, E, L>(v, sm, t, l);
  // ---- pass 1 ----
  pass_compute<N, E, S::R1, S::R0>(v, t, a.tw, tw_unit);
  constexpr bool three = S::R2 > 1;
  if constexpr (three) {
    __syncthreads();
    pass_to_smem<N, E, S::R1, S::R0, L>(v, sm, t, l);
    __syncthreads();
    smem_to_regs<N, E, L>(v, sm, t, l);
    pass_compute<N, E, (three ? S::R2 : 2), S::R0 * S::R1>(v, t, a.tw, tw_unit);
  }
  constexpr int RL = three ? S::R2 : S::R1;            // radix of the last pass
  constexpr int NSL = N / RL;                          // its NS

  if (MODE == FFT_R2C_FWD) {
    // packed Z -> X_m, m = 0..N (external/ffte-7.0/dzfft2d.f NY=1 branch == rfft), times scale (= 1/np)
    __syncthreads();
    pass_to_smem<N, E, RL, NSL, L>(v, sm, t, l);
    __syncthreads();
#pragma unroll
    for (int j = 0; j <= E; ++j) {
      if (j == E && t != 0) break;
      const int m = (j == E) ? N : t + T * j;
      const int m1 = (m == N) ? 0 : m;
      const int m2 = (m == 0) ? 0 : N - m;
      cplx zm = sm[m1 * L + l];
      cplx zc = cconj(sm[m2 * L + l]);
      cplx e = make_double2(0.5 * (zm.x + zc.x), 0.5 * (zm.y + zc.y));
      cplx d = csub(zm, zc);
      cplx o = make_double2(0.5 * d.y, -0.5 * d.x);   // (-i/2) d
      cplx x = cadd(e, cmul(__ldg(&a.tw[m]), o));
      if (ok) {
        const cplx val = make_double2(x.x * a.scale, x.y * a.scale);
        if (EXT && a.use_peer) {
          const long long kk = q / a.batch0;                 // z plane
          const int ii = (int)(q - kk * a.batch0);           // local row
          int dq;
          size_t dst;
          slab_put_index(0, a.pt.rank, a.pt.nranks, a.pt.r_cnt, a.pt.r_off, a.pt.m_cnt, a.pt.m_off, a.nrdim, 0, ii, m,
                         (int)kk, &dq, &dst);
          reinterpret_cast<cplx *>(reinterpret_cast<char *>(a.pt.base[dq]) + a.pt.data_off + blockIdx.y * a.pt.fstride)[dst] =
              val;
        } else {
          gout[(long long)m * a.stride_pt] = val;
        }
      }
    }
    return;
  }

  // ---- last pass straight to HBM ----
  constexpr int QL = E / RL;
#pragma unroll
  for (int qq = 0; qq < QL; ++qq)
#pragma unroll
    for (int u = 0; u < RL; ++u) {
      cplx x = v[qq + u * QL];
      if (MODE == FFT_C2C_BWD || MODE == FFT_C2R_BWD) x = cconj(x);
      const int idx = out_index<N, E, RL, NSL>(t, qq, u);
      if (ok) {
        cplx val = make_double2(x.x * a.scale, x.y * a.scale);
        if (EXT && MODE == FFT_C2R_BWD && scaled) val = make_double2(div_by(val.x, rv, rinv), div_by(val.y, rv, rinv));
        gout[(long long)idx * a.stride_pt] = val;
      }
    }
  if (MODE == FFT_C2R_BWD) {
    // padding column keeps the Nyquist input times np (quirk Q3; ops:1702-1705); a separate rscale pass would divide it too
    const double fac = (double)(2 * N);
    if (t == 0 && ok) {
      cplx val = make_double2(nyq.x * fac, nyq.y * fac);
      if (EXT && scaled) val = make_double2(div_by(val.x, rv, rinv), div_by(val.y, rv, rinv));
      gout[(long long)N * a.stride_pt] = val;
    }
  }
}

// resident CTAs per SM the kernels are compiled for: 5 (4 for c2r) for the plain short-line kernels (48 / 64 registers),
// 2 for the 16-points-per-thread ones (an explicit 1 lets ptxas take 140-172 registers and halves the occupancy)
template <int MODE, int N, int E, int THREADS, int EXT>
struct MinBlocks {
  static constexpr int value =
      THREADS != 256 ? 1 : (E <= 8 ? (EXT ? 3 : (MODE == FFT_C2R_BWD ? 4 : 5)) : 2);
};

template <int MODE, int N, int E, int THREADS, int EXT>
__global__ void __launch_bounds__(THREADS, MinBlocks<MODE, N, E, THREADS, EXT>::value) fft_reg_kernel(FftRegArgs a) {
  if (!EXT) {
    fft_reg_tile<MODE, N, E, THREADS, 0>(a, blockIdx.x);   // one tile per CTA
    return;
  }
  constexpr int L = THREADS / (N / E);
  const long long ntiles = (a.nlines + L - 1) / L;
  // the fused exchange runs a grid-stride loop so that the system-scope fence that ends it (an NVLink round trip during
  // which the CTA still holds its SM resources) is paid once per CTA, not once per tile
  for (long long blk = blockIdx.x; blk < ntiles; blk += gridDim.x) {
    fft_reg_tile<MODE, N, E, THREADS, EXT>(a, blk);
    if (blk + gridDim.x < ntiles) __syncthreads();   // the tile's shared buffer is reused
  }
  if (MODE == FFT_R2C_FWD && a.use_peer) dist_finish_put(a.pt, gridDim.x * gridDim.y);
}

// ---- host side ---------------------------------------------------------------------------------

template <int N, int E, int THREADS>
struct RegCfg {
  static constexpr int T = N / E, L = THREADS / T;
  static constexpr size_t smem = (size_t)(N + 1) * L * sizeof(cplx);
};

template <int MODE, int N, int E, int THREADS, int EXT>
static int launch_one_ext(const FftRegArgs &a, int nfields, cudaStream_t st) {
  using C = RegCfg<N, E, THREADS>;
  static bool attr = false;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(fft_reg_kernel<MODE, N, E, THREADS, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)C::smem));
    attr = true;
  }
  long long ntiles = (a.nlines + C::L - 1) / C::L;
  if (a.use_peer) {
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      CUDA_TRY(cudaGetDevice(&dev));
      CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int per_sm = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft_reg_kernel<MODE, N, E, THREADS, EXT>, THREADS,
                                                          C::smem));
    const long long resident = (long long)sms * std::max(per_sm, 1);
    ntiles = std::min(ntiles, std::max(1ll, resident / nfields));
  }
  const dim3 grid((unsigned)ntiles, (unsigned)nfields);
  fft_reg_kernel<MODE, N, E, THREADS, EXT><<<grid, THREADS, C::smem, st>>>(a);
  return MLEGS_OK;
}

template <int MODE, int N, int E, int THREADS>
static int launch_one(const FftRegArgs &a, int nfields, cudaStream_t st) {
  // the extended kernel only exists for the real modes (fused exchange: r2c; fused row scaling: r2c and c2r)
  if constexpr (MODE == FFT_R2C_FWD || MODE == FFT_C2R_BWD) {
    if (MODE == FFT_C2R_BWD && a.rs.perm_p > 1) return launch_one_ext<MODE, N, E, THREADS, 2>(a, nfields, st);
    if (a.use_peer || a.rs.mode != 0) return launch_one_ext<MODE, N, E, THREADS, 1>(a, nfields, st);
  }
  return launch_one_ext<MODE, N, E, THREADS, 0>(a, nfields, st);
}

template <int MODE>
static int launch_mode(int n, const FftRegArgs &a, int nfields, cudaStream_t st) {
  switch (n) {
    case 32: return launch_one<MODE, 32, 8, 256>(a, nfields, st);
    // measured at 128^3 (profiles/r2/README.md): 8 points per thread and 3 passes for N = 128: 51 vs 46 us; 128-thread CTAs
    // (16 lines each, twice as many per SM) for N = 64 / 128: 58 / 47 vs 60 / 47 us -- occupancy is not what holds these
    // kernels at 0.66-0.74 of the copy bandwidth
    case 64: return launch_one<MODE, 64, 8, 256>(a, nfields, st);
    case 128: return launch_one<MODE, 128, 16, 256>(a, nfields, st);
    case 256: return launch_one<MODE, 256, 16, 256>(a, nfields, st);
    case 512: return launch_one<MODE, 512, 16, 256>(a, nfields, st);   // 8 lines per CTA, 3 CTAs per SM: 0.87 -> 0.78 ms at 512^3
    case 1024: return launch_one<MODE, 1024, 16, 512>(a, nfields, st);   // (4 lines per CTA, 3 CTAs per SM: 482 vs 489 us)
  }
  return fail(MLEGS_E_ARG, "fft_reg: unsupported length");
}

bool fft_reg_supported(int n) { return n == 32 || n == 64 || n == 128 || n == 256 || n == 512 || n == 1024; }

int launch_fft_reg(FftMode mode, int n, const cplx *in, cplx *out, long long nlines, long long batch0,
                   long long stride_b1, long long stride_pt, const double *tw, int tw_order, double scale,
                   const int *colstart, int ncols, int nrl, cudaStream_t st, const PeerTable *peer, int nrdim,
                   const FieldBatch *fb, const RowScale *rs) {
  FftRegArgs a;
  if (rs) a.rs = *rs;
  int nfields = 1;
  if (fb && fb->n > 0) {
    nfields = fb->n;
    for (int i = 0; i < nfields; ++i) {
      a.in[i] = fb->in[i];
      a.out[i] = fb->out[i];
    }
  } else {
    a.in[0] = in;
    a.out[0] = out;
  }
  static const char *dbg = getenv("MLEGS_FFT_COPY");
  a.dbg_copy = dbg && dbg[0] == '1';
  a.use_peer = peer != nullptr;
  a.nrdim = nrdim;
  if (peer) a.pt = *peer;
  a.nlines = nlines;
  a.batch0 = batch0;
  a.stride_b1 = stride_b1;
  a.stride_pt = stride_pt;
  a.tw = reinterpret_cast<const cplx *>(tw);
  a.tw_order = tw_order;
  a.scale = scale;
  a.colstart = colstart;
  a.ncols = ncols;
  a.nrl = nrl;
  switch (mode) {
    case FFT_C2C_FWD: return launch_mode<FFT_C2C_FWD>(n, a, nfields, st);
    case FFT_C2C_BWD: return launch_mode<FFT_C2C_BWD>(n, a, nfields, st);
    case FFT_R2C_FWD: return launch_mode<FFT_R2C_FWD>(n, a, nfields, st);
    case FFT_C2R_BWD: return launch_mode<FFT_C2R_BWD>(n, a, nfields, st);
  }
  return MLEGS_OK;
}

}  // namespace mlegs
