// Radial mapped-Legendre transform as batched FP64 tensor-core GEMMs (DMMA m8n8k4), one GEMM
// problem per azimuthal wavenumber m.  Replaces rtrans_forward / rtrans_backward of
// /root/reference/src/submodules/mlegs_scalar_ops.f90:1852-2008, which promote the real table to
// complex and call zgemm once per m on strided slices.
//
// The table pf(i, n, m) is real, the data complex: the complex columns are treated as 2*nz real
// columns, so the contraction is a real GEMM (half the flops of the reference's zgemm).  Parity
// folding (f(i) +- f(nr+1-i)) halves the contraction length again.
//
//   forward  (analysis):  a(n,k) = sum_{i<nr/2} pf(i,n,m) * w(i) * (f(i,k) + (-1)^n f(nr-1-i,k))
//   backward (synthesis): be(i,k) = sum_{n even} pf(i,n,m) a(n,k),  bo likewise over odd n,
//                         f(i) = be + bo,  f(nr-1-i) = be - bo
//
// A CTA owns one (m, 32 complex columns) strip and a 128-row (forward: n, both parities) or
// 64-row (backward: i, both parity accumulators) output tile; the contraction runs in chunks of 16
// through a two-stage cp.async (LDGSTS) pipeline: the table slice and the RAW field data of chunk c+1
// stream into shared memory while the tensor pipe works on chunk c, and the parity fold (f(i) +- f(nr-1-i))
// times the quadrature weight is applied when the B fragments are read.  8 warps: 2 parities x (2 x 2)
// warp tiles of 32 x 32, 16 DMMA tiles per warp per k-step.
#include "kernels.h"

namespace mlegs {

#define LEG_THREADS 256
#define LEG_KC 16          // contraction chunk
#define LEG_NTC 32         // complex columns per CTA  (64 real columns)
#define LEG_LD (LEG_KC + 4)  // padded leading dimension of K-contiguous smem tiles (conflict-free frags)
#define LEG_MT_F 128       // forward: consecutive n per CTA (64 even + 64 odd)
#define LEG_MT_B 64        // backward: i rows per CTA
#define LEG_LDA_B (LEG_MT_B + 4)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ int nn_of_m(int mglob, int nrc, int npc) {
  if (mglob >= npc) return 0;
  int v = min(nrc, nrc - mglob);
  return v > 0 ? v : 0;
}

// ------------------------------------------------------------------------------------------------
// asynchronous global -> shared copies (LDGSTS); src_bytes == 0 zero-fills the destination
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int nbytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int nbytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(nbytes) : "memory");
}
// DRAM -> L2 prefetch of a contiguous range (bytes: multiple of 16, 16-byte aligned address); no smem involved
__device__ __forceinline__ void l2_prefetch(const void *gmem, int bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gmem), "r"(bytes) : "memory");
}
#define LEG_PD 3   // chunks of field data kept in flight towards L2 ahead of the cp.async stage

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
#define LEG_LDB_F (2 * LEG_KC + 8)       // doubles per kz row of the raw (unfolded) data tiles: conflict-free fragments
struct FwdSmem {
  double A[2][LEG_MT_F / 2][LEG_LD];     // [parity][row within parity][k]
  double T[LEG_NTC][LEG_LDB_F];          // raw top rows    f(i, kz), i = chunk rows, (re, im) interleaved
  double Bm[LEG_NTC][LEG_LDB_F];         // raw mirror rows f(nr-1-i, kz)
  double W[LEG_KC];                      // quadrature weights of the chunk rows
};

__global__ void __launch_bounds__(LEG_THREADS, 2) leg_forward_kernel(LegArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  FwdSmem *sm = reinterpret_cast<FwdSmem *>(smraw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int par = warp >> 2, wr = (warp >> 1) & 1, wc = warp & 1;
  const int ml = blockIdx.z;
  const int mglob = a.m0 + ml;
  const int nn = (a.skip_m0 && mglob == 0) ? 0 : nn_of_m(mglob, a.nrc, a.npc);
  const int n0 = blockIdx.y * LEG_MT_F;
  const int kz0 = blockIdx.x * LEG_NTC;
  const size_t col_stride = (size_t)a.nrl * a.npl;     // elements between z planes
  const cplx *in = a.in + (size_t)ml * a.nrl;
  cplx *out = a.out + (size_t)ml * a.nrl;
  const double *pf = a.pf + (size_t)mglob * a.nrh * a.ne;
  const bool use_ln = (mglob == 0) && (a.lnval != 0.0);
  const bool have_w = a.w != nullptr;
  const bool vec2 = (a.nrh & 1) == 0;                  // table rows are 16-byte aligned

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const bool active = (n0 < nn);
  const int nchunks = active ? (a.nrh + LEG_KC - 1) / LEG_KC : 0;

  // Stage chunk c (rows i = c*KC .. +KC of the half grid) into buffer `buf`; out-of-range pieces are zero-filled.
  auto issue = [&](int c, int buf) {
    FwdSmem &S = sm[buf];
    const int i0 = c * LEG_KC;
    // A: 128 table rows x KC doubles
    if (vec2) {
      // 128 x 8 16-byte pieces -> 4 per thread
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = tid + LEG_THREADS * j;
        const int r = p >> 3, k2 = (p & 7) * 2;
        const int n = n0 + r;
        const bool ok = (n < nn) && (i0 + k2 < a.nrh);
        cp_async16(&S.A[r & 1][r >> 1][k2], ok ? (const void *)&pf[(size_t)n * a.nrh + i0 + k2] : (const void *)pf, ok);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int p = tid + LEG_THREADS * j;
        const int r = p >> 4, k = p & 15;
        const int n = n0 + r;
        const bool ok = (n < nn) && (i0 + k < a.nrh);
        cp_async8(&S.A[r & 1][r >> 1][k], ok ? (const void *)&pf[(size_t)n * a.nrh + i0 + k] : (const void *)pf, ok);
      }
    }
    // raw data: 32 kz x KC complex, top and mirrored rows -> 2 + 2 per thread
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int p = tid + LEG_THREADS * j;
      const int kzl = p >> 4, k = p & 15;
      const int kz = kz0 + kzl, i = i0 + k;
      const bool ok = (kz < a.nzl) && (i < a.nrh);
      const cplx *col = in + (size_t)kz * col_stride;
      cp_async16(&S.T[kzl][2 * k], ok ? (const void *)&col[i] : (const void *)in, ok);
      cp_async16(&S.Bm[kzl][2 * k], ok ? (const void *)&col[a.nr - 1 - i] : (const void *)in, ok);
    }
    if (have_w && tid < LEG_KC) {
      const bool ok = i0 + tid < a.nrh;
      cp_async8(&S.W[tid], ok ? (const void *)&a.w[i0 + tid] : (const void *)a.w, ok);
    }
  };

  // The field columns of this CTA are streamed exactly once and come from DRAM in 256-byte pieces: keep LEG_PD
  // chunks of them on their way into L2 so that the cp.async stage only ever sees L2 latency.
  auto prefetch = [&](int c) {
    if (tid < 2 * LEG_NTC) {
      const int kz = kz0 + (tid & (LEG_NTC - 1)), i0 = c * LEG_KC;
      if (c < nchunks && kz < a.nzl && i0 < a.nrh) {
        const int cnt = min(LEG_KC, a.nrh - i0);
        const cplx *col = in + (size_t)kz * col_stride;
        l2_prefetch((tid >> 5) ? (const void *)&col[a.nr - i0 - cnt] : (const void *)&col[i0], cnt * 16);
      }
    }
  };

  const int fr = lane >> 2, fk = lane & 3;
  // The warp's four 8-row tiles are interleaved with the other row-warp's (tile index 2 mt + wr), and tiles
  // that lie entirely beyond the truncation nn(m) are skipped: both row-warps (hence all four SM sub-partitions)
  // keep the same number of DMMAs when nn is not a multiple of the 128-row CTA tile.
  int nact = 0;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) nact += (n0 + 2 * ((2 * mt + wr) * 8) + par < nn) ? 1 : 0;
  const bool minus = (par ^ a.swap_parity) != 0;        // fold sign: even rows take f(i) + f(mirror), odd rows the difference
  const int bre = fr & 1;                               // this lane's real column is the re (0) or im (1) part

  if (nchunks > 0) {
    issue(0, 0);
    cp_async_commit();
#pragma unroll
    for (int d = 1; d <= LEG_PD; ++d) prefetch(d);
  }
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    cp_async_wait_all();
    __syncthreads();            // chunk c has landed for everyone; everyone is done reading buffer buf^1
    if (c + 1 < nchunks) {
      issue(c + 1, buf ^ 1);
      cp_async_commit();
      prefetch(c + 1 + LEG_PD);
    }
    const FwdSmem &S = sm[buf];
#pragma unroll
    for (int ks = 0; ks < LEG_KC / 4; ++ks) {
      const int k = ks * 4 + fk;
      double af[4], bf[4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) af[mt] = S.A[par][(2 * mt + wr) * 8 + fr][k];
      double wk = have_w ? S.W[k] : 1.0;
      double l1 = 0.0, l2 = 0.0;
      if (use_ln && bre == 0) {   // log term removed from the real part of the m = 0 column (ops:193-195)
        const int i = c * LEG_KC + k;
        if (i < a.nrh) {
          l1 = a.lnval * __ldg(&a.lnx[i]);
          l2 = a.lnval * __ldg(&a.lnx[a.nr - 1 - i]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int kzl = (wc * 32 + nt * 8 + fr) >> 1;
        double t = S.T[kzl][2 * k + bre];
        double b = S.Bm[kzl][2 * k + bre];
        if (use_ln) {
          t -= l1;
          b -= l2;
        }
        double f = minus ? (t - b) : (t + b);
        bf[nt] = have_w ? f * wk : f;
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
        if (mt < nact) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
    }
  }

  // epilogue: thread holds C[row = tile*8 + lane/4][cols 2*(lane%4), +1] of each 8x8 tile == one complex.
  // The two parities interleave along n, so the tile goes through shared memory and leaves as contiguous
  // 16-byte-per-lane rows (n fastest), 2 KB per kz column.
  __syncthreads();            // everyone is done with the staging buffers
  cplx(*cs)[LEG_MT_F + 2] = reinterpret_cast<cplx(*)[LEG_MT_F + 2]>(smraw);
  static_assert(sizeof(cplx) * LEG_NTC * (LEG_MT_F + 2) <= 2 * sizeof(FwdSmem), "epilogue smem");
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const int nl = 2 * ((2 * mt + wr) * 8 + fr) + par;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int kzl = (wc * 32 + nt * 8) / 2 + fk;
      cs[kzl][nl] = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < LEG_MT_F * LEG_NTC / LEG_THREADS; ++j) {
    const int p = tid + LEG_THREADS * j;
    const int nl = p & (LEG_MT_F - 1), kzl = p / LEG_MT_F;
    const int n = n0 + nl, kz = kz0 + kzl;
    if (n < a.nrdim && kz < a.nzl) out[(size_t)kz * col_stride + n] = cs[kzl][nl];
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
#define LEG_LDB_B (4 * LEG_KC + 2)       // doubles per kz row of the raw coefficient tile (32 complex + pad)
struct BwdSmem {
  double A[2][LEG_KC][LEG_LDA_B];        // [parity][k (coefficient pair index)][i]
  double B[LEG_NTC][LEG_LDB_B];          // raw coefficients a(n, kz), n = 32 consecutive, (re, im) interleaved
};

__global__ void __launch_bounds__(LEG_THREADS, 2) leg_backward_kernel(LegArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  BwdSmem *sm = reinterpret_cast<BwdSmem *>(smraw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int par = warp >> 2, wr = (warp >> 1) & 1, wc = warp & 1;
  const int ml = blockIdx.z;
  const int mglob = a.m0 + ml;
  const int nn = nn_of_m(mglob, a.nrc, a.npc);
  const int i0 = blockIdx.y * LEG_MT_B;
  const int kz0 = blockIdx.x * LEG_NTC;
  const size_t col_stride = (size_t)a.nrl * a.npl;
  const cplx *in = a.in + (size_t)ml * a.nrl;
  cplx *out = a.out + (size_t)ml * a.nrl;
  const double *pf = a.pf + (size_t)mglob * a.nrh * a.ne;
  const bool use_ln = (mglob == 0) && (a.lnval != 0.0);
  const bool vec2 = (a.nrh & 1) == 0;

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int kpairs = (nn + 1) / 2;                       // contraction length of the even parity (>= odd)
  const int nchunks = (kpairs + LEG_KC - 1) / LEG_KC;

  auto issue = [&](int c, int buf) {
    BwdSmem &S = sm[buf];
    const int nbase = c * 2 * LEG_KC;                    // first coefficient index n of this chunk
    // A: 32 consecutive n x 64 i doubles
    if (vec2) {
      // 32 x 32 16-byte pieces -> 4 per thread
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = tid + LEG_THREADS * j;
        const int nl = p >> 5, i2 = (p & 31) * 2;
        const int n = nbase + nl, ii = i0 + i2;
        const bool ok = (n < nn) && (ii < a.nrh);
        cp_async16(&S.A[nl & 1][nl >> 1][i2], ok ? (const void *)&pf[(size_t)n * a.nrh + ii] : (const void *)pf, ok);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int p = tid + LEG_THREADS * j;
        const int nl = p >> 6, il = p & 63;
        const int n = nbase + nl, ii = i0 + il;
        const bool ok = (n < nn) && (ii < a.nrh);
        cp_async8(&S.A[nl & 1][nl >> 1][il], ok ? (const void *)&pf[(size_t)n * a.nrh + ii] : (const void *)pf, ok);
      }
    }
    // raw coefficients: 32 kz x 32 n complex -> 4 per thread
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = tid + LEG_THREADS * j;
      const int kzl = p >> 5, nl = p & 31;
      const int kz = kz0 + kzl, n = nbase + nl;
      const bool ok = (kz < a.nzl) && (n < nn);
      cp_async16(&S.B[kzl][2 * nl], ok ? (const void *)&in[(size_t)kz * col_stride + n] : (const void *)in, ok);
    }
  };

  auto prefetch = [&](int c) {   // DRAM -> L2 for the coefficient columns, LEG_PD chunks ahead (512-byte pieces)
    if (tid < LEG_NTC) {
      const int kz = kz0 + tid, nb = c * 2 * LEG_KC;
      if (c < nchunks && kz < a.nzl && nb < nn) {
        const int cnt = min(2 * LEG_KC, nn - nb);
        l2_prefetch(&in[(size_t)kz * col_stride + nb], cnt * 16);
      }
    }
  };
  const int fr = lane >> 2, fk = lane & 3;
  const int bre = fr & 1;
  if (nchunks > 0) {
    issue(0, 0);
    cp_async_commit();
#pragma unroll
    for (int d = 1; d <= LEG_PD; ++d) prefetch(d);
  }
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    cp_async_wait_all();
    __syncthreads();
    if (c + 1 < nchunks) {
      issue(c + 1, buf ^ 1);
      cp_async_commit();
      prefetch(c + 1 + LEG_PD);
    }
    const BwdSmem &S = sm[buf];
#pragma unroll
    for (int ks = 0; ks < LEG_KC / 4; ++ks) {
      const int k = ks * 4 + fk;
      double af[4], bf[4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) af[mt] = S.A[par][k][wr * 32 + mt * 8 + fr];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bf[nt] = S.B[(wc * 32 + nt * 8 + fr) >> 1][2 * (2 * k + par) + bre];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
  }
  __syncthreads();   // everyone is done with the staging buffers before they are reused below

  // combine parities through shared memory: C[par][i (64)][real col (64)], padded
  double(*cs)[LEG_MT_B][2 * LEG_NTC + 2] = reinterpret_cast<double(*)[LEG_MT_B][2 * LEG_NTC + 2]>(smraw);
  static_assert(sizeof(double) * 2 * LEG_MT_B * (2 * LEG_NTC + 2) <= sizeof(BwdSmem) * 2, "epilogue smem");
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      int r = wr * 32 + mt * 8 + fr;
      int cc = wc * 32 + nt * 8 + 2 * fk;
      cs[par][r][cc] = acc[mt][nt][0];
      cs[par][r][cc + 1] = acc[mt][nt][1];
    }
  __syncthreads();
  // 64 i x 32 kz outputs (x2 mirrored); thread -> (i = tid & 63, kz = tid>>6 + 4 j)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int i = tid & 63;
    int kzl = (tid >> 6) + 4 * j;
    int ii = i0 + i, kz = kz0 + kzl;
    if (ii < a.nrh && kz < a.nzl) {
      double er = cs[0][i][2 * kzl], ei = cs[0][i][2 * kzl + 1];
      double orr = cs[1][i][2 * kzl], oi = cs[1][i][2 * kzl + 1];
      cplx top = make_double2(er + orr, ei + oi);
      cplx bot = make_double2(er - orr, ei - oi);
      if (use_ln) {
        top.x += a.lnval * __ldg(&a.lnx[ii]);
        bot.x += a.lnval * __ldg(&a.lnx[a.nr - 1 - ii]);
      }
      out[(size_t)kz * col_stride + ii] = top;
      out[(size_t)kz * col_stride + (a.nr - 1 - ii)] = bot;
    }
  }
  // rows nr .. nrdim-1 are zero after rtrans_backward (se = 0 initialisation, ops:1975-1976)
  if (blockIdx.y == 0) {
    int npad = a.nrdim - a.nr;
    for (int idx = tid; idx < npad * LEG_NTC; idx += LEG_THREADS) {
      int r = idx % npad, kz = kz0 + idx / npad;
      if (kz < a.nzl) out[(size_t)kz * col_stride + a.nr + r] = make_double2(0.0, 0.0);
    }
  }
}

int setup_leg_kernels() {
  CUDA_TRY(cudaFuncSetAttribute(leg_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(2 * sizeof(FwdSmem))));
  CUDA_TRY(cudaFuncSetAttribute(leg_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(2 * sizeof(BwdSmem))));
  return MLEGS_OK;
}

int launch_leg_forward(const LegArgs &a, cudaStream_t st) {
  if (a.npl <= 0 || a.nzl <= 0) return MLEGS_OK;
  dim3 grid((a.nzl + LEG_NTC - 1) / LEG_NTC, (a.nrdim + LEG_MT_F - 1) / LEG_MT_F, a.npl);
  prof_begin("legendre_forward", st);
  leg_forward_kernel<<<grid, LEG_THREADS, 2 * sizeof(FwdSmem), st>>>(a);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

int launch_leg_backward(const LegArgs &a, cudaStream_t st) {
  if (a.npl <= 0 || a.nzl <= 0) return MLEGS_OK;
  dim3 grid((a.nzl + LEG_NTC - 1) / LEG_NTC, (a.nrh + LEG_MT_B - 1) / LEG_MT_B, a.npl);
  prof_begin("legendre_backward", st);
  leg_backward_kernel<<<grid, LEG_THREADS, 2 * sizeof(BwdSmem), st>>>(a);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

}  // namespace mlegs
