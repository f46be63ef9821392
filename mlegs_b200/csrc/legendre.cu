// Radial mapped-Legendre transform as batched FP64 tensor-core GEMMs (DMMA m8n8k4), one GEMM
// problem per azimuthal wavenumber m.  These are the cp.async kernels of the first half of round 1; the production
// path is the warp-specialised TMA-fed pair in legendre_ws.cu (launch_leg_* below dispatch to it), and the kernels
// here remain for radial sizes whose table rows are not 16-byte aligned (nr not a multiple of 4) and for A/B timing
// (MLEGS_LEG_NO_WS=1).  Replaces rtrans_forward / rtrans_backward of
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
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "dist_dev.cuh"

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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// Per CTA: A (table slice) and the RAW field rows are double-buffered cp.async targets; a cooperative fold pass
// (each element once per CTA: (f(i) +- f(nr-1-i)) * w(i), 16 FP64 operations per thread and chunk) turns the raw
// rows of chunk c+1 into the two parity operands F while the tensor pipe works on chunk c.  Keeping the fold out
// of the fragment path matters: plain FP64 instructions share the pipe with DMMA and would otherwise sit in
// front of every k-step of every warp.
struct FwdSmem {
  double A[2][2][LEG_MT_F / 2][LEG_LD];     // [buf][parity][row within parity][k]
  double F[2][2][2 * LEG_NTC][LEG_LD];      // [buf][fold: 0 sum, 1 difference][real column][k]
  cplx T[2][LEG_NTC][LEG_KC];               // [buf] raw top rows    f(i, kz)
  cplx Bm[2][LEG_NTC][LEG_KC];              // [buf] raw mirror rows f(nr-1-i, kz)
  double W[2][LEG_KC];                      // [buf] quadrature weights of the chunk rows
};

template <int NACT>
struct IntC {
  static constexpr int value = NACT;
};

__global__ void __launch_bounds__(LEG_THREADS, 2) leg_forward_kernel(LegArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  FwdSmem &S = *reinterpret_cast<FwdSmem *>(smraw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int par = warp >> 2, wr = (warp >> 1) & 1, wc = warp & 1;
  // scalars of one launch are neighbours in the grid so that they share the table slice of their m in L2
  const int fld = blockIdx.z % a.fb.n, ml = blockIdx.z / a.fb.n;
  const int mglob = a.m0 + ml * a.ms;
  const int nn = (a.skip_m0 && mglob == 0) ? 0 : nn_of_m(mglob, a.nrc, a.npc);
  const int n0 = blockIdx.y * LEG_MT_F;
  const int kz0 = blockIdx.x * LEG_NTC;
  const size_t col_stride = (size_t)a.nrl * a.npl;     // elements between z planes
  const cplx *in = a.fb.in[fld] + (size_t)ml * a.nrl;
  cplx *out = a.fb.out[fld] + (size_t)ml * a.nrl;
  const double *pf = a.pf + (size_t)mglob * a.nrh * a.ne;
  const double lnval = a.fb.ln[fld];
  const bool use_ln = (mglob == 0) && (lnval != 0.0);
  const bool have_w = a.w != nullptr;
  const bool vec2 = (a.nrh & 1) == 0;                  // table rows are 16-byte aligned

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const bool active = (n0 < nn);
  const int nchunks = active ? (a.nrh + LEG_KC - 1) / LEG_KC : 0;

  // table rows of chunk c -> A[buf]; out-of-range pieces are zero-filled
  auto issue_A = [&](int c, int buf) {
    const int i0 = c * LEG_KC;
    if (vec2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {          // 128 rows x 8 16-byte pieces
        const int p = tid + LEG_THREADS * j;
        const int r = p >> 3, k2 = (p & 7) * 2;
        const int n = n0 + r;
        const bool ok = (n < nn) && (i0 + k2 < a.nrh);
        cp_async16(&S.A[buf][r & 1][r >> 1][k2], ok ? (const void *)&pf[(size_t)n * a.nrh + i0 + k2] : (const void *)pf,
                   ok);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int p = tid + LEG_THREADS * j;
        const int r = p >> 4, k = p & 15;
        const int n = n0 + r;
        const bool ok = (n < nn) && (i0 + k < a.nrh);
        cp_async8(&S.A[buf][r & 1][r >> 1][k], ok ? (const void *)&pf[(size_t)n * a.nrh + i0 + k] : (const void *)pf, ok);
      }
    }
  };
  // raw field rows (top and mirror) + weights of chunk c -> T/Bm/W[buf]
  auto issue_raw = [&](int c, int buf) {
    const int i0 = c * LEG_KC;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int p = tid + LEG_THREADS * j;
      const int kzl = p >> 4, k = p & 15;
      const int kz = kz0 + kzl, i = i0 + k;
      const bool ok = (kz < a.nzl) && (i < a.nrh);
      const cplx *col = in + (size_t)kz * col_stride;
      cp_async16(&S.T[buf][kzl][k], ok ? (const void *)&col[i] : (const void *)in, ok);
      cp_async16(&S.Bm[buf][kzl][k], ok ? (const void *)&col[a.nr - 1 - i] : (const void *)in, ok);
    }
    if (have_w && tid < LEG_KC) {
      const bool ok = i0 + tid < a.nrh;
      cp_async8(&S.W[buf][tid], ok ? (const void *)&a.w[i0 + tid] : (const void *)a.w, ok);
    }
  };
  // fold raw chunk c (in T/Bm/W[rbuf]) into the parity operands F[fbuf]; thread -> (k = tid & 15, kz = tid>>4 + 16 j)
  auto fold = [&](int c, int rbuf, int fbuf) {
    const int k = tid & 15;
    const double wk = have_w ? S.W[rbuf][k] : 1.0;
    double l1 = 0.0, l2 = 0.0;
    if (use_ln) {   // log term removed from the real part of the m = 0 column (ops:193-195)
      const int i = c * LEG_KC + k;
      if (i < a.nrh) {
        l1 = lnval * __ldg(&a.lnx[i]);
        l2 = lnval * __ldg(&a.lnx[a.nr - 1 - i]);
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int kzl = (tid >> 4) + 16 * j;
      cplx t = S.T[rbuf][kzl][k], b = S.Bm[rbuf][kzl][k];
      t.x -= l1;
      b.x -= l2;
      S.F[fbuf][0][2 * kzl][k] = (t.x + b.x) * wk;
      S.F[fbuf][0][2 * kzl + 1][k] = (t.y + b.y) * wk;
      S.F[fbuf][1][2 * kzl][k] = (t.x - b.x) * wk;
      S.F[fbuf][1][2 * kzl + 1][k] = (t.y - b.y) * wk;
    }
  };

  const int fr = lane >> 2, fk = lane & 3;
  // The warp's four 8-row tiles are interleaved with the other row-warp's (tile index 2 mt + wr); tiles entirely
  // beyond the truncation nn(m) are skipped by compiling the main loop for each possible count (a predicated-off
  // DMMA still occupies the pipe).
  int nact = 0;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) nact += (n0 + 2 * ((2 * mt + wr) * 8) + par < nn) ? 1 : 0;
  const int fsel = par ^ a.swap_parity;                 // even rows contract with the sum fold, odd rows with the difference

  auto mainloop = [&](auto nact_c) {
    constexpr int NACT = decltype(nact_c)::value;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      cp_async_wait_all();
      __syncthreads();     // A(c), raw(c+1) have landed; F(c) is complete; everyone has left iteration c-1
      if (c + 1 < nchunks) issue_A(c + 1, buf ^ 1);
      if (c + 2 < nchunks) issue_raw(c + 2, buf);
      cp_async_commit();
      if (c + 1 < nchunks) fold(c + 1, buf ^ 1, buf ^ 1);
#pragma unroll
      for (int ks = 0; ks < LEG_KC / 4; ++ks) {
        const int k = ks * 4 + fk;
        double af[4], bf[4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
          if (mt < NACT) af[mt] = S.A[buf][par][(2 * mt + wr) * 8 + fr][k];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) bf[nt] = S.F[buf][fsel][wc * 32 + nt * 8 + fr][k];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
          if (mt < NACT) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
          }
      }
    }
  };

  if (nchunks > 0) {
    issue_A(0, 0);
    issue_raw(0, 0);
    cp_async_commit();
    if (nchunks > 1) issue_raw(1, 1);       // both raw chunks travel together: one DRAM latency, not two
    cp_async_commit();
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    __syncthreads();
    fold(0, 0, 0);
    switch (nact) {
      case 4: mainloop(IntC<4>{}); break;
      case 3: mainloop(IntC<3>{}); break;
      case 2: mainloop(IntC<2>{}); break;
      case 1: mainloop(IntC<1>{}); break;
      default: mainloop(IntC<0>{}); break;
    }
  }

  // epilogue: thread holds C[row = tile*8 + lane/4][cols 2*(lane%4), +1] of each 8x8 tile == one complex.
  // The two parities interleave along n, so the tile goes through shared memory and leaves as contiguous
  // 16-byte-per-lane rows (n fastest), 2 KB per kz column.
  __syncthreads();            // everyone is done with the staging buffers
  cplx(*cs)[LEG_MT_F + 2] = reinterpret_cast<cplx(*)[LEG_MT_F + 2]>(smraw);
  static_assert(sizeof(cplx) * LEG_NTC * (LEG_MT_F + 2) <= sizeof(FwdSmem), "epilogue smem");
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
  // the grid only covers rows below the largest nn(m); the last row tile zeroes the rest (se = 0, ops:1898-1899)
  if (blockIdx.y == gridDim.y - 1) {
    const int nfirst = n0 + LEG_MT_F, nrest = a.nrdim - nfirst;
    for (int p = tid; p < nrest * LEG_NTC; p += LEG_THREADS) {
      const int n = nfirst + p % nrest, kz = kz0 + p / nrest;
      if (kz < a.nzl) out[(size_t)kz * col_stride + n] = make_double2(0.0, 0.0);
    }
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

// PUT: the physical-space rows are stored straight into the windows of the ranks that own them (exchange(1,2) fused
// into the epilogue) and the kernel ends with the exchange barrier.
template <bool PUT>
__global__ void __launch_bounds__(LEG_THREADS, 2) leg_backward_kernel(LegArgs a, PeerTable pt) {
  extern __shared__ __align__(16) unsigned char smraw[];
  BwdSmem *sm = reinterpret_cast<BwdSmem *>(smraw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int par = warp >> 2, wr = (warp >> 1) & 1, wc = warp & 1;
  const int fld = blockIdx.z % a.fb.n, ml = blockIdx.z / a.fb.n;
  const int mglob = a.m0 + ml * a.ms;
  const int nn = nn_of_m(mglob, a.nrc, a.npc);
  const int i0 = blockIdx.y * LEG_MT_B;
  const int kz0 = blockIdx.x * LEG_NTC;
  const size_t col_stride = (size_t)a.nrl * a.npl;
  const cplx *in = a.fb.in[fld] + (size_t)ml * a.nrl;
  cplx *out = a.fb.out[fld] + (size_t)ml * a.nrl;
  const double *pf = a.pf + (size_t)mglob * a.nrh * a.ne;
  const double lnval = a.fb.ln[fld];
  const bool use_ln = (mglob == 0) && (lnval != 0.0);
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

  const int fr = lane >> 2, fk = lane & 3;
  const int bre = fr & 1;
  if (nchunks > 0) {
    issue(0, 0);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    cp_async_wait_all();
    __syncthreads();
    if (c + 1 < nchunks) {
      issue(c + 1, buf ^ 1);
      cp_async_commit();
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
        top.x += lnval * __ldg(&a.lnx[ii]);
        bot.x += lnval * __ldg(&a.lnx[a.nr - 1 - ii]);
      }
      if (PUT) {
        int dq;
        size_t dst;
        slab_put_index(1, pt.rank, pt.nranks, pt.r_cnt, pt.r_off, pt.m_cnt, pt.m_off, a.nrdim, a.npdim, ii, ml, kz, &dq,
                       &dst);
        reinterpret_cast<cplx *>(reinterpret_cast<char *>(pt.base[dq]) + pt.data_off)[dst] = top;
        slab_put_index(1, pt.rank, pt.nranks, pt.r_cnt, pt.r_off, pt.m_cnt, pt.m_off, a.nrdim, a.npdim, a.nr - 1 - ii, ml,
                       kz, &dq, &dst);
        reinterpret_cast<cplx *>(reinterpret_cast<char *>(pt.base[dq]) + pt.data_off)[dst] = bot;
      } else {
        out[(size_t)kz * col_stride + ii] = top;
        out[(size_t)kz * col_stride + (a.nr - 1 - ii)] = bot;
      }
    }
  }
  // rows nr .. nrdim-1 are zero after rtrans_backward (se = 0 initialisation, ops:1975-1976)
  if (blockIdx.y == 0) {
    int npad = a.nrdim - a.nr;
    for (int idx = tid; idx < npad * LEG_NTC; idx += LEG_THREADS) {
      int r = idx % npad, kz = kz0 + idx / npad;
      if (kz < a.nzl) {
        if (PUT) {
          int dq;
          size_t dst;
          slab_put_index(1, pt.rank, pt.nranks, pt.r_cnt, pt.r_off, pt.m_cnt, pt.m_off, a.nrdim, a.npdim, a.nr + r, ml, kz,
                         &dq, &dst);
          reinterpret_cast<cplx *>(reinterpret_cast<char *>(pt.base[dq]) + pt.data_off)[dst] = make_double2(0.0, 0.0);
        } else {
          out[(size_t)kz * col_stride + a.nr + r] = make_double2(0.0, 0.0);
        }
      }
    }
  }
  if (PUT) dist_finish_put(pt, gridDim.x * gridDim.y * gridDim.z);
}

int setup_leg_kernels() {
  CUDA_TRY(cudaFuncSetAttribute(leg_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(FwdSmem)));
  CUDA_TRY(cudaFuncSetAttribute(leg_backward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(2 * sizeof(BwdSmem))));
  CUDA_TRY(cudaFuncSetAttribute(leg_backward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(2 * sizeof(BwdSmem))));
  return MLEGS_OK;
}

// single-scalar callers leave fb.n == 0: the launch then handles in/out/lnval
static LegArgs with_batch(const LegArgs &a0) {
  LegArgs a = a0;
  if (a.fb.n <= 0) {
    a.fb.n = 1;
    a.fb.in[0] = a.in;
    a.fb.out[0] = a.out;
    a.fb.ln[0] = a.lnval;
  }
  return a;
}

bool leg_ws_forward_supported(const LegArgs &a);              // legendre_ws.cu
int launch_leg_forward_ws(const LegArgs &a, cudaStream_t st);

int launch_leg_forward(const LegArgs &a0, cudaStream_t st) {
  const LegArgs a = with_batch(a0);
  if (a.npl <= 0 || a.nzl <= 0) return MLEGS_OK;
  static const bool no_ws = getenv("MLEGS_LEG_NO_WS") != nullptr;   // A/B timing against the cp.async kernel
  if (!no_ws && leg_ws_forward_supported(a)) return launch_leg_forward_ws(a, st);
  // row tiles: up to the largest truncation among the local columns (the last tile zero-fills rows beyond it)
  int nn_max = (a.m0 < a.npc) ? std::max(std::min(a.nrc, a.nrc - a.m0), 0) : 0;
  nn_max = std::max(1, std::min(nn_max, a.nrdim));
  dim3 grid((a.nzl + LEG_NTC - 1) / LEG_NTC, (nn_max + LEG_MT_F - 1) / LEG_MT_F, a.npl * a.fb.n);
  prof_begin("legendre_forward", st);
  leg_forward_kernel<<<grid, LEG_THREADS, sizeof(FwdSmem), st>>>(a);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

bool leg_ws_supported(const LegArgs &a);                      // legendre_ws.cu
int launch_leg_backward_ws(const LegArgs &a, cudaStream_t st);

int launch_leg_backward(const LegArgs &a0, cudaStream_t st) {
  const LegArgs a = with_batch(a0);
  if (a.npl <= 0 || a.nzl <= 0) return MLEGS_OK;
  static const bool no_ws = getenv("MLEGS_LEG_NO_WS") != nullptr;   // A/B timing against the cp.async kernel
  if (!no_ws && leg_ws_supported(a)) return launch_leg_backward_ws(a, st);
  if (a.peer && a.fb.n != 1) return fail(MLEGS_E_STATE, "rtrans_backward: the fused exchange handles one scalar per launch");
  dim3 grid((a.nzl + LEG_NTC - 1) / LEG_NTC, (a.nrh + LEG_MT_B - 1) / LEG_MT_B, a.npl * a.fb.n);
  prof_begin(a.peer ? "legendre_backward_put" : "legendre_backward", st);
  if (a.peer) {
    leg_backward_kernel<true><<<grid, LEG_THREADS, 2 * sizeof(BwdSmem), st>>>(a, *a.peer);
  } else {
    PeerTable none;
    memset(&none, 0, sizeof(none));
    leg_backward_kernel<false><<<grid, LEG_THREADS, 2 * sizeof(BwdSmem), st>>>(a, none);
  }
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

}  // namespace mlegs
