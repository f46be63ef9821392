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
// through shared memory with register-staged prefetch of the next chunk.  8 warps: 2 parities x
// (2 x 2) warp tiles of 32 x 32, 16 DMMA tiles per warp per k-step.
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
// forward
// ------------------------------------------------------------------------------------------------
struct FwdSmem {
  double A[2][LEG_MT_F / 2][LEG_LD];     // [parity][row within parity][k]
  double B[2][2 * LEG_NTC][LEG_LD];      // [fold: 0 even(+), 1 odd(-)][real column][k]
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

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const bool active = (n0 < nn);
  const int nchunks = active ? (a.nrh + LEG_KC - 1) / LEG_KC : 0;

  // staging registers: A: 128 rows x 16 k = 2048 doubles -> 8 per thread; thread -> (k = tid & 15, rows tid>>4 + 16 j)
  // B: 32 kz x 16 i -> 512 (top,bottom) complex pairs -> 2 per thread; thread -> (i = tid & 15, kz = tid>>4 + 16 j)
  double ra[8];
  cplx rbe[2], rbo[2];
  const int lk = tid & 15, lr = tid >> 4;

  auto gload = [&](int c) {
    const int kk = c * LEG_KC + lk;     // i index
    const bool kok = kk < a.nrh;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + lr + 16 * j;
      ra[j] = (kok && n < nn) ? __ldg(&pf[(size_t)n * a.nrh + kk]) : 0.0;
    }
    double wi = kok ? (a.w ? __ldg(&a.w[kk]) : 1.0) : 0.0;
    double l1 = 0.0, l2 = 0.0;
    if (use_ln && kok) {
      l1 = a.lnval * __ldg(&a.lnx[kk]);
      l2 = a.lnval * __ldg(&a.lnx[a.nr - 1 - kk]);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int kz = kz0 + lr + 16 * j;
      cplx top = make_double2(0.0, 0.0), bot = make_double2(0.0, 0.0);
      if (kok && kz < a.nzl) {
        top = in[(size_t)kz * col_stride + kk];
        bot = in[(size_t)kz * col_stride + (a.nr - 1 - kk)];
        top.x -= l1;
        bot.x -= l2;
      }
      rbe[j] = make_double2((top.x + bot.x) * wi, (top.y + bot.y) * wi);
      rbo[j] = make_double2((top.x - bot.x) * wi, (top.y - bot.y) * wi);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int r = lr + 16 * j;             // row within tile; n0 is even so parity(r) == parity(n)
      sm[buf].A[r & 1][r >> 1][lk] = ra[j];
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int c = 2 * (lr + 16 * j);
      sm[buf].B[0][c][lk] = rbe[j].x;
      sm[buf].B[0][c + 1][lk] = rbe[j].y;
      sm[buf].B[1][c][lk] = rbo[j].x;
      sm[buf].B[1][c + 1][lk] = rbo[j].y;
    }
  };

  if (nchunks > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  const int fr = lane >> 2, fk = lane & 3;
  // The warp's four 8-row tiles are interleaved with the other row-warp's (tile index 2 mt + wr), and tiles
  // that lie entirely beyond the truncation nn(m) are skipped: both row-warps (hence all four SM sub-partitions)
  // keep the same number of DMMAs when nn is not a multiple of the 128-row CTA tile.
  int nact = 0;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) nact += (n0 + 2 * ((2 * mt + wr) * 8) + par < nn) ? 1 : 0;
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) gload(c + 1);
#pragma unroll
    for (int ks = 0; ks < LEG_KC / 4; ++ks) {
      double af[4], bf[4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) af[mt] = sm[buf].A[par][(2 * mt + wr) * 8 + fr][ks * 4 + fk];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bf[nt] = sm[buf].B[par ^ a.swap_parity][wc * 32 + nt * 8 + fr][ks * 4 + fk];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
        if (mt < nact) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
    }
    if (c + 1 < nchunks) sstore(buf ^ 1);
    __syncthreads();
  }

  // epilogue: thread holds C[row = tile*8 + lane/4][cols 2*(lane%4), +1] of each 8x8 tile == one complex
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    int n = n0 + 2 * ((2 * mt + wr) * 8 + fr) + par;
    if (n >= a.nrdim) continue;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      int kz = kz0 + (wc * 32 + nt * 8) / 2 + fk;
      if (kz < a.nzl) out[(size_t)kz * col_stride + n] = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct BwdSmem {
  double A[2][LEG_KC][LEG_LDA_B];        // [parity][k (coefficient pair index)][i]
  double B[2][2 * LEG_NTC][LEG_LD];      // [parity][real column][k]
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

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int kpairs = (nn + 1) / 2;                       // contraction length of the even parity (>= odd)
  const int nchunks = (kpairs + LEG_KC - 1) / LEG_KC;

  // A staging: 2 parities x 16 k x 64 i = 2048 doubles -> 8 per thread; thread -> (i = tid & 63, q = tid >> 6 (0..3))
  //            element j: kk2 = q + 4 j (0..31) -> (parity = kk2 & 1, k = kk2 >> 1): consecutive n
  // B staging: 32 consecutive n x 32 kz complex = 1024 -> 4 per thread; thread -> (nloc = tid & 31, kz = tid>>5 + 8 j)
  double ra[8];
  cplx rb[4];
  const int li = tid & 63, lq = tid >> 6;
  const int lnl = tid & 31, lkz = tid >> 5;

  auto gload = [&](int c) {
    const int nbase = c * 2 * LEG_KC;                    // first coefficient index n of this chunk
    const int ii = i0 + li;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = nbase + lq + 4 * j;
      ra[j] = (ii < a.nrh && n < nn) ? __ldg(&pf[(size_t)n * a.nrh + ii]) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = nbase + lnl;
      int kz = kz0 + lkz + 8 * j;
      rb[j] = (n < nn && kz < a.nzl) ? in[(size_t)kz * col_stride + n] : make_double2(0.0, 0.0);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int k2 = lq + 4 * j;
      sm[buf].A[k2 & 1][k2 >> 1][li] = ra[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = 2 * (lkz + 8 * j);
      sm[buf].B[lnl & 1][c][lnl >> 1] = rb[j].x;
      sm[buf].B[lnl & 1][c + 1][lnl >> 1] = rb[j].y;
    }
  };

  if (nchunks > 0) {
    gload(0);
    sstore(0);
  }
  __syncthreads();
  const int fr = lane >> 2, fk = lane & 3;
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) gload(c + 1);
#pragma unroll
    for (int ks = 0; ks < LEG_KC / 4; ++ks) {
      double af[4], bf[4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) af[mt] = sm[buf].A[par][ks * 4 + fk][wr * 32 + mt * 8 + fr];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bf[nt] = sm[buf].B[par][wc * 32 + nt * 8 + fr][ks * 4 + fk];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
    if (c + 1 < nchunks) sstore(buf ^ 1);
    __syncthreads();
  }

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
