// Warp-specialised, TMA-fed (cp.async.bulk + mbarrier) persistent kernels of the radial mapped-Legendre transform:
// the production path behind launch_leg_backward / launch_leg_forward (legendre.cu keeps the cp.async kernels for
// shapes the bulk copies cannot address).  Replaces rtrans_backward / rtrans_forward of
// /root/reference/src/submodules/mlegs_scalar_ops.f90:1852-2008.
//
// Why this shape (profiles/r1, profiles/r2 and tools/microbench): one DMMA-issuing warp per SM sub-partition can
// saturate the FP64 tensor pipe in a pure DMMA stream (dmma_shapes.cu: 37.1 TFLOP/s with 4 warps per SM), but a kernel
// whose warps also compute addresses, issue cp.async, wait for their own loads, meet at a CTA barrier every 16
// contraction steps and stage the epilogue through shared memory left the pipe idle half of the time.  So:
//   * producer warps walk a list of output tiles (heaviest first; a static snake schedule for short tiles, a work
//     counter for long ones) and move whole rows of the table and of the field with cp.async.bulk / TMA tensor boxes
//     (no per-thread address arithmetic, no register staging), signalling full[stage] through the mbarrier
//     transaction count;
//   * consumer (DMMA) warps wait on full[stage], run the stage's DMMAs, arrive on empty[stage], and store their
//     accumulators straight to HBM at the end of a tile.  Every warp accumulates BOTH parities of its rows, so the
//     parity combination f(i) = be + bo, f(nr-1-i) = be - bo happens in registers.  No CTA-wide barrier after start-up,
//     and no prologue: rows and columns that no tile computes are zero-filled by the DMMA warps next to their tiles.
// Two tile shapes (template parameter NTC = z planes per tile):
//   * NTC = 32: 8 DMMA + 4 producer warps, one persistent CTA per SM, deep stage ring;
//   * NTC = 16: 4 DMMA + 2 producer warps, TWO persistent CTAs per SM (half the table reuse, but two independent
//     pipelines per SM whose epilogues and stalls interleave); used by the synthesis of short transforms (nr <= 128).
// What was tried and measured at 128^3 (8 scalars per launch; in-kernel globaltimer timelines, profiles/r2/README.md):
// the r1 kernel spent 13 of 95 us in a zero-fill prologue (removed: 87 us); a work counter shared by 148 CTAs costs
// ~2 us per fetch but was hidden behind the stage ring (static schedule: no change); forcing the two row groups of a CTA
// to alternate on the tensor pipe made it worse (100 us: one warp per sub-partition reaches 58 % of the pipe in this
// loop, two reach 83 %); an epilogue staged through the vacated stage slot and sent by TMA bulk stores was slower than
// the direct stores (103 us).
#include <cuda.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>

#include "dist_dev.cuh"

namespace mlegs {

// NTC z planes per tile -> NTC/8 column groups of DMMA warps x 2 row groups, one producer warp per 8 planes (a bulk
// copy is a warp-uniform instruction, UBLKCP, so the copies of one warp issue one after the other)
template <int NTC, int HW = 2>
struct WsCfg {
  static constexpr int NCG = NTC / 8;          // column groups (8 z planes = 16 real columns each)
  static constexpr int NCW = HW * NCG;         // DMMA warps: HW row groups x NCG column groups
  static constexpr int NPW = NCG;              // producer warps
  static constexpr int CONS = 32 * NCW, PROD = 32 * NPW, THREADS = CONS + PROD;
};
#define WS_RING 8                    // decoded tiles in flight
#define WS_KC 16                     // contraction steps per stage

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MB_DONE;\n"
      "bra MB_WAIT;\n"
      "MB_DONE:\n"
      "}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (TMA unit, SASS UBLKCP); bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ int ws_nn_of_m(int mglob, int nrc, int npc) {
  if (mglob >= npc) return 0;
  int v = min(nrc, nrc - mglob);
  return v > 0 ? v : 0;
}

template <int N>
struct IntK {
  static constexpr int value = N;
};

// ------------------------------------------------------------------------------------------------
// backward (synthesis):  be(i,k) = sum_{n even} pf(i,n,m) a(n,k),  bo likewise over odd n,
//                        f(i) = be + bo,  f(nr-1-i) = be - bo
// ------------------------------------------------------------------------------------------------
// Output tile = 64 rows i x NTC z planes of one column m of one scalar; a stage = 32 consecutive n (16 of each parity):
// the table rows pf(i0 .. i0+63, n, m) (512 contiguous bytes each) and the coefficient rows a(n .. n+31, kz) of the
// tile's planes (512 contiguous bytes each).
#define WSB_MT 64
#define WSB_LDA (WSB_MT + 4)            // doubles per table row in shared memory   (fragment reads conflict-free)
#define WSB_LDB (4 * WS_KC + 2)         // doubles per coefficient row               (idem)

template <int NTC>
struct BwdStageWS {
  double A[2][WS_KC][WSB_LDA];          // [parity of n][pair index][i]
  double B[NTC][WSB_LDB];               // [kz][n (32 consecutive), (re, im) interleaved]
};
struct BwdCtxWS {
  const double *pf;                     // table slice of this column, at row i0
  const cplx *in;                       // coefficients a(0, m, kz0)
  cplx *out;
  double lnval;
  int nn, i0, kz0, ml, fld, valid;
};
template <int NTC, int NS>
struct BwdSmemWS {
  BwdStageWS<NTC> st[NS];
  BwdCtxWS ring[WS_RING];
  unsigned long long full[NS], empty[NS];
};
static_assert(sizeof(BwdSmemWS<32, 6>) <= 227 * 1024, "shared memory of the backward kernel");
static_assert(2 * (sizeof(BwdSmemWS<16, 4>) + 1024) <= 227 * 1024, "two backward CTAs per SM");
static_assert((WSB_LDA * 8) % 16 == 0 && (WSB_LDB * 8) % 16 == 0, "bulk copy destinations must be 16-byte aligned");

struct LegItemsB {
  int total, nkz, nfld, nit, mcount;    // tiles = mcount columns x nfld scalars x nit row tiles x nkz plane tiles
};

// Static tile schedule: round j of the (heaviest-first) tile list goes to the CTAs in snake order -- CTA b takes tile
// j G + b in even rounds and j G + (G-1-b) in odd ones -- so at any moment the grid works on G consecutive tiles (the L2
// sharing the tile order is built for) and every CTA gets the same mix of heavy and light tiles.
__device__ __forceinline__ int ws_next_tile(int &round) {
  const int G = (int)gridDim.x, b = (int)blockIdx.x;
  const int it = round * G + ((round & 1) ? G - 1 - b : b);
  ++round;
  return it;
}

template <int NTC>
__device__ __forceinline__ void bwd_fetch(const LegArgs &a, const LegItemsB &L, BwdCtxWS *cx, int &round) {
  const int it = ws_next_tile(round);
  if (it >= L.total) {
    cx->valid = 0;
    return;
  }
  // kz tile fastest, then row tile, then scalar, then column: concurrent tiles share the table slice of their m in L2
  const int kzt = it % L.nkz;
  int rest = it / L.nkz;
  const int itile = rest % L.nit;
  rest /= L.nit;
  const int fld = rest % L.nfld, ml = rest / L.nfld;
  const int mglob = a.m0 + ml * a.ms;
  cx->i0 = itile * WSB_MT;
  cx->kz0 = kzt * NTC;
  cx->pf = a.pf + (size_t)mglob * a.nrh * a.ne + cx->i0;
  cx->in = a.fb.in[fld] + (size_t)ml * a.nrl + (size_t)cx->kz0 * a.nrl * a.npl;
  cx->out = a.fb.out[fld] + (size_t)ml * a.nrl;
  cx->lnval = (mglob == 0) ? a.fb.ln[fld] : 0.0;
  cx->nn = ws_nn_of_m(mglob, a.nrc, a.npc);
  cx->ml = ml;
  cx->fld = fld;
  cx->valid = 1;
}

// MODE 0: rows stored in place (i, m, kz).  MODE 1 (several ranks): rows stored straight into the windows of the ranks
// that own them in physical space (exchange(1,2) fused into the epilogue) + the exchange barrier.  MODE 2: rows stored
// into the LOCAL output buffer in destination order (slab_stage_index); dist.cu's ship kernel then moves them as long
// contiguous runs -- at 8 ranks the direct puts are 128-byte pieces and reach a third of the NVLink rate.
template <int MODE>
__device__ __forceinline__ void bwd_store(const LegArgs &a, const PeerTable &pt, int fld, int ml, int kz, int row,
                                          size_t col_stride, cplx v) {
  if (MODE == 1) {
    int dq;
    size_t dst;
    slab_put_index(1, pt.rank, pt.nranks, pt.r_cnt, pt.r_off, pt.m_cnt, pt.m_off, a.nrdim, a.npdim, row, ml, kz, &dq, &dst);
    reinterpret_cast<cplx *>(reinterpret_cast<char *>(pt.base[dq]) + pt.data_off + fld * pt.fstride)[dst] = v;
  } else if (MODE == 2) {
    a.fb.out[fld][slab_stage_index(pt.rank, pt.nranks, pt.r_cnt, pt.r_off, pt.m_cnt, a.nzl, row, ml, kz)] = v;
  } else {
    a.fb.out[fld][(size_t)kz * col_stride + (size_t)ml * a.nrl + row] = v;
  }
}

template <int MODE, int NTC, int NS, int MINB>
__global__ void __launch_bounds__(WsCfg<NTC>::THREADS, MINB) leg_backward_ws_kernel(LegArgs a, LegItemsB L, PeerTable pt) {
  using Cfg = WsCfg<NTC>;
  constexpr int CONS = Cfg::CONS, PROD = Cfg::PROD, NCG = Cfg::NCG;
  extern __shared__ __align__(128) unsigned char smraw[];
  BwdSmemWS<NTC, NS> &S = *reinterpret_cast<BwdSmemWS<NTC, NS> *>(smraw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t col_stride = (size_t)a.nrl * a.npl;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      mbar_init(&S.full[s], PROD);
      mbar_init(&S.empty[s], Cfg::NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // The producer publishes tile jj+1 in the ring before it fills any stage of tile jj, so a consumer that has waited
  // for the stages of tile jj may read ring[jj+1]; tiles 0 and 1 are published before the start-up barrier.  A
  // consumer therefore never waits on a stage of a tile that does not exist.
  int round = 0;                        // position in this CTA's tile schedule (used by the first producer thread)
  if (tid == CONS) {
    bwd_fetch<NTC>(a, L, &S.ring[0], round);
    bwd_fetch<NTC>(a, L, &S.ring[1], round);
  }
  __syncthreads();   // barriers initialised, first tiles published

  if (warp >= Cfg::NCW) {
    // =============================== producer warps ===============================
    // warp pw moves table rows TR pw .. TR pw + TR-1 (lanes 0 .. TR-1) and coefficient rows 8 pw .. 8 pw + 7 (the next
    // 8 lanes) of a stage
    constexpr int TR = 2 * WS_KC / Cfg::NPW;
    const int pw = warp - Cfg::NCW;
    const int arow_l = (lane < TR) ? pw * TR + lane : -1;
    const int brow_l = (lane >= TR && lane < TR + 8) ? pw * 8 + lane - TR : -1;
    int fetched = 2;
    auto ensure = [&](int j) {          // ring entries 0 .. j have been written (by the first producer thread)
      while (fetched <= j) {
        if (tid == CONS) bwd_fetch<NTC>(a, L, &S.ring[fetched & (WS_RING - 1)], round);
        ++fetched;
        asm volatile("bar.sync 1, %0;\n" ::"n"(PROD) : "memory");
      }
    };
    int g = 0;                          // stage counter over all tiles of this CTA
    for (int jj = 0;; ++jj) {
      ensure(jj + 1);
      const BwdCtxWS x = S.ring[jj & (WS_RING - 1)];
      if (!x.valid) break;
      const int nch = ((x.nn + 1) / 2 + WS_KC - 1) / WS_KC;
      const unsigned arow = (unsigned)min(WSB_MT, a.nrh - x.i0) * 8u;       // bytes of one table row piece
      const int kz = x.kz0 + brow_l;
      for (int c = 0; c < nch; ++c, ++g) {
        const int s = g % NS;
        if (g >= NS) mbar_wait(&S.empty[s], ((g / NS) + 1) & 1);            // consumers are done with stage use g - NS
        BwdStageWS<NTC> &B = S.st[s];
        const int nbase = c * 2 * WS_KC;
        const int n = nbase + arow_l;                                       // this lane's table row
        const int nvalid = min(2 * WS_KC, x.nn - nbase);                    // coefficient rows present in this stage
        const bool do_a = arow_l >= 0 && n < x.nn;
        const bool do_b = brow_l >= 0 && kz < a.nzl;
        unsigned bytes = 0;
        if (do_a) bytes += arow;
        if (do_b) bytes += (unsigned)nvalid * 16u;
        // The consumers run the k-steps up to the last retained pair, rounded up to 4 pairs: table rows in that
        // range beyond nn(m) are not copied and must not hold stale NaN patterns (they multiply zero coefficients).
        if (arow_l >= 0 && n >= x.nn && n < nbase + 8 * ((nvalid + 7) >> 3)) {
          double2 *row = reinterpret_cast<double2 *>(&B.A[n & 1][(n - nbase) >> 1][0]);
          for (int q = 0; q < WSB_MT / 2; ++q) row[q] = make_double2(0.0, 0.0);
        }
        if (nvalid < 2 * WS_KC && do_b) {
          // coefficients beyond the truncation are not stored anywhere (the axial FFT only carries the retained
          // rows): they are zeros of the contraction.  Table rows beyond nn(m) then multiply zeros.
          for (int q = nvalid; q < 2 * WS_KC; ++q) {
            B.B[brow_l][2 * q] = 0.0;
            B.B[brow_l][2 * q + 1] = 0.0;
          }
        }
        mbar_arrive_expect_tx(&S.full[s], bytes);
        if (do_a) bulk_g2s(&B.A[n & 1][(n - nbase) >> 1][0], x.pf + (size_t)n * a.nrh, arow, &S.full[s]);
        if (do_b) bulk_g2s(&B.B[brow_l][0], x.in + (size_t)brow_l * col_stride + nbase, (unsigned)nvalid * 16u, &S.full[s]);
      }
    }
  } else {
    // =============================== consumer warps ===============================
    // warp (h, wq): rows i0 + 32 h .. + 31 (4 tiles of 8), real columns 16 wq .. 16 wq + 15 (2 tiles of 8), both parities
    const int h = warp / NCG, wq = warp % NCG;
    const int fr = lane >> 2, fk = lane & 3;
    double acc[2][4][2][2];

    // Whole columns without a retained coefficient are zeros (se = 0, ops:1975-1976); they have no tile.  Written here,
    // while the first stages are in flight.  (The padding rows nr .. nrdim-1 of the other columns are written by the
    // epilogue of their first row tile.)
    {
      const int nzero = a.npl - L.mcount;
      const int ncol = L.nfld * nzero * a.nzl;
      for (int col = blockIdx.x * Cfg::NCW + warp; col < ncol; col += gridDim.x * Cfg::NCW) {
        const int kz = col % a.nzl, rest = col / a.nzl;
        const int ml = L.mcount + rest % nzero, fld = rest / nzero;
        for (int r = lane; r < a.nrdim; r += 32) bwd_store<MODE>(a, pt, fld, ml, kz, r, col_stride, make_double2(0.0, 0.0));
      }
    }

    int g = 0;
    for (int jj = 0;; ++jj) {
      const BwdCtxWS &x = S.ring[jj & (WS_RING - 1)];
      if (!x.valid) break;
      const int nn = x.nn, i0 = x.i0, kz0 = x.kz0, ml = x.ml, fld = x.fld;
      const double lnval = x.lnval;
      const int kpairs = (nn + 1) / 2;
      const int nch = (kpairs + WS_KC - 1) / WS_KC;
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) acc[p][i][j][0] = acc[p][i][j][1] = 0.0;

      for (int c = 0; c < nch; ++c, ++g) {
        const int s = g % NS;
        mbar_wait(&S.full[s], (g / NS) & 1);
        const BwdStageWS<NTC> &B = S.st[s];
        const int ksteps = min(WS_KC / 4, (kpairs - c * WS_KC + 3) >> 2);     // pairs beyond the truncation are zeros
#pragma unroll
        for (int ks = 0; ks < WS_KC / 4; ++ks) {
          if (ks >= ksteps) break;
          const int k = ks * 4 + fk;
          double af[2][4], bf[2][2];
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) af[p][mt] = B.A[p][k][h * 32 + mt * 8 + fr];
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) bf[p][nt] = B.B[(wq * 16 + nt * 8 + fr) >> 1][2 * (2 * k + p) + (fr & 1)];
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) dmma884(acc[p][mt][nt][0], acc[p][mt][nt][1], af[p][mt], bf[p][nt]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.empty[s]);
      }

      // epilogue: thread holds (re, im) of be and bo at row i = i0 + 32 h + 8 mt + lane/4, plane kz0 + 8 wq + 4 nt + lane%4
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const int ii = i0 + h * 32 + mt * 8 + fr;
        double l1 = 0.0, l2 = 0.0;
        if (__double_as_longlong(lnval) != 0 && ii < a.nrh) {   // + ln term of the m = 0 column (ops:219-221)
          l1 = lnval * __ldg(&a.lnx[ii]);
          l2 = lnval * __ldg(&a.lnx[a.nr - 1 - ii]);
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int kz = kz0 + wq * 8 + nt * 4 + fk;
          if (ii < a.nrh && kz < a.nzl) {
            const double er = acc[0][mt][nt][0], ei = acc[0][mt][nt][1];
            const double orr = acc[1][mt][nt][0], oi = acc[1][mt][nt][1];
            bwd_store<MODE>(a, pt, fld, ml, kz, ii, col_stride, make_double2(er + orr + l1, ei + oi));
            bwd_store<MODE>(a, pt, fld, ml, kz, a.nr - 1 - ii, col_stride, make_double2(er - orr + l2, ei - oi));
          }
        }
      }
      // first row tile of its column: the padding rows nr .. nrdim-1 are zeros (se = 0, ops:1975-1976)
      if (i0 == 0) {
        const int kz = kz0 + wq * 8 + (lane >> 2);
        if (kz < a.nzl)
          for (int r = a.nr + 4 * h + (lane & 3); r < a.nrdim; r += 8)
            bwd_store<MODE>(a, pt, fld, ml, kz, r, col_stride, make_double2(0.0, 0.0));
      }
    }
  }   // consumers
  if (MODE == 1) dist_finish_put(pt, gridDim.x);
}

// ------------------------------------------------------------------------------------------------
// forward (analysis):  a(n,k) = sum_{i<nr/2} [pf(i,n,m) w(i)] * (f(i,k) + (-1)^n f(nr-1-i,k))
// ------------------------------------------------------------------------------------------------
// Output tile = 128 consecutive n (64 of each parity) x NTC z planes of one column m of one scalar; a stage = 32
// radial points i.  The table slice arrives as two TMA tensor boxes (16 i x 128 n, 128-byte swizzle: fragment reads are
// bank-conflict free without padding); the field rows f(i0.., kz) and their mirrors f(.. nr-1-i0, kz) arrive as 512-byte
// bulk copies.  The quadrature weight is folded into the table once at start-up (pf*w, resident in HBM), so the parity
// fold left in the DMMA warps' fragment path is one DADD per B fragment: (top + mirror) feeds the rows of one
// parity, (top - mirror) the other.
#define WSF_MT 128
// KC = radial points per stage (16 or 32); a field row in shared memory is KC + 4 complex elements long (fragment reads
// conflict-free for KC = 16 and 32: the row stride is 8 mod 16 doubles)

template <int NTC, int KC>
struct FwdStageWS {
  double A[KC / 16][WSF_MT][16];         // swizzled TMA boxes: [k / 16][n][k % 16, 16-byte chunks XOR (n & 7)]
  double T[NTC][2 * (KC + 4)];           // f(i0 + k, kz), (re, im) interleaved
  double Bm[NTC][2 * (KC + 4)];          // f(nr-1-i0-k, kz) at complex index KC-1-k
};
struct FwdCtxWS {
  const cplx *in;                        // f(0, m, kz0)
  cplx *out;
  double lnval;
  int nn, n0, kz0, mglob, valid;
};
template <int NTC, int KC, int NS>
struct FwdSmemWS {
  FwdStageWS<NTC, KC> st[NS];
  FwdCtxWS ring[WS_RING];
  unsigned long long full[NS], empty[NS];
};
static_assert(sizeof(FwdSmemWS<32, 32, 3>) <= 227 * 1024, "shared memory of the forward kernel");
static_assert(sizeof(FwdSmemWS<32, 16, 6>) <= 227 * 1024, "shared memory of the forward kernel (16-point stages)");
static_assert(2 * (sizeof(FwdSmemWS<16, 32, 2>) + 1024) <= 227 * 1024, "two forward CTAs per SM");
static_assert(sizeof(FwdStageWS<32, 32>) % 1024 == 0 && sizeof(FwdStageWS<16, 32>) % 1024 == 0 &&
                  sizeof(FwdStageWS<32, 16>) % 1024 == 0,
              "swizzled boxes need 1024-byte alignment");

// Tile order: row tile fastest, then z tile, then scalar, then column m (heaviest columns first).  The row tiles of one
// (m, scalar, z tile) read the same field rows and run at the same time on neighbouring SMs, so the field is fetched
// from HBM once and the other row tiles hit in L2 (at 512^3 the row-tile-major order re-read it 4x: ncu 4.0 GB
// against 1.1 GB); the table slice of (m, row tile) is shared by the z tiles and scalars that follow within the
// same few hundred tiles.  Row tiles beyond the truncation of their column are skipped at fetch time.
struct LegItemsF {
  int total, nkz, nfld, mlo, nrt;        // nrt = row tiles of the widest column
  int mcount;                            // columns [mlo, mlo + mcount) have retained rows
  unsigned int *ctr;                     // != nullptr: tiles are claimed from this counter instead of the static schedule
};

template <int NTC>
__device__ __forceinline__ void fwd_fetch(const LegArgs &a, const LegItemsF &L, FwdCtxWS *cx, int &round) {
  for (;;) {
    // Long tiles (nr >= 512: 8+ stages each) are claimed from a counter: the row tiles of one (m, z tile) then run at
    // the same time and share their field rows in L2 (a static schedule drifts apart: 512^3 forward 3.5 -> 4.1 ms).
    // Short tiles use the static schedule: a fetch from a counter every CTA hits costs more than their DMMA work.
    const int it = L.ctr ? (int)atomicAdd(L.ctr, 1u) : ws_next_tile(round);
    if (it >= L.total) {
      cx->valid = 0;
      return;
    }
    const int r = it % L.nrt;
    int rest = it / L.nrt;
    const int kzt = rest % L.nkz;
    rest /= L.nkz;
    const int fld = rest % L.nfld, ml = L.mlo + rest / L.nfld;
    const int mglob = a.m0 + ml * a.ms;
    const int nn = ws_nn_of_m(mglob, a.nrc, a.npc);
    if (r * WSF_MT >= nn) continue;      // nothing retained in this row tile: the last row tile with work zero-fills it
    cx->kz0 = kzt * NTC;
    cx->in = a.fb.in[fld] + (size_t)ml * a.nrl + (size_t)cx->kz0 * a.nrl * a.npl;
    cx->out = a.fb.out[fld] + (size_t)ml * a.nrl;
    cx->lnval = (mglob == 0) ? a.fb.ln[fld] : 0.0;
    cx->nn = nn;
    cx->n0 = r * WSF_MT;
    cx->mglob = mglob;
    cx->valid = 1;
    return;
  }
}

// 3-D TMA box load: table(i, n, m) -> swizzled shared box
__device__ __forceinline__ void tma_box3(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// HW = row groups of DMMA warps (2: 64 rows per warp, 4: 32 rows per warp and four DMMA warps per SM sub-partition)
template <int NTC, int KC, int NS, int MINB, int HW, int CREGS>
__global__ void __launch_bounds__(WsCfg<NTC, HW>::THREADS, MINB)
    leg_forward_ws_kernel(LegArgs a, LegItemsF L, const __grid_constant__ CUtensorMap tmap) {
  using Cfg = WsCfg<NTC, HW>;
  constexpr int MT = 8 / HW;            // 8-row DMMA tiles of each parity per warp
  constexpr int CONS = Cfg::CONS, PROD = Cfg::PROD, NCG = Cfg::NCG;
  extern __shared__ __align__(1024) unsigned char smraw_f[];
  FwdSmemWS<NTC, KC, NS> &S = *reinterpret_cast<FwdSmemWS<NTC, KC, NS> *>(smraw_f);   // 1024-aligned
  if ((smem_u32(smraw_f) & 1023u) != 0) __trap();                              // the swizzled boxes rely on it
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t col_stride = (size_t)a.nrl * a.npl;
  const int NCH = (a.nrh + KC - 1) / KC;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      mbar_init(&S.full[s], PROD);
      mbar_init(&S.empty[s], Cfg::NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  int round = 0;                        // position in this CTA's tile schedule (used by the first producer thread)
  if (tid == CONS) {
    fwd_fetch<NTC>(a, L, &S.ring[0], round);
    fwd_fetch<NTC>(a, L, &S.ring[1], round);
  }
  __syncthreads();   // barriers initialised, first tiles published (same protocol as the backward kernel)

  if (warp >= Cfg::NCW) {
    // =============================== producer warps ===============================
    // warp pw: field rows (z planes) 8 pw .. 8 pw + 7: lanes 0-7 the top rows, lanes 8-15 the mirror rows; the first
    // producer thread also issues the two table boxes
    if (CREGS > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    const int pw = warp - Cfg::NCW;
    const int row_l = pw * 8 + (lane & 7);
    const bool is_top = lane < 8, is_mir = lane >= 8 && lane < 16;
    int fetched = 2;
    auto ensure = [&](int j) {
      while (fetched <= j) {
        if (tid == CONS) fwd_fetch<NTC>(a, L, &S.ring[fetched & (WS_RING - 1)], round);
        ++fetched;
        asm volatile("bar.sync 1, %0;\n" ::"n"(PROD) : "memory");
      }
    };
    int g = 0;
    for (int jj = 0;; ++jj) {
      ensure(jj + 1);
      const FwdCtxWS x = S.ring[jj & (WS_RING - 1)];
      if (!x.valid) break;
      const bool row_ok = (is_top || is_mir) && (x.kz0 + row_l < a.nzl);
      const cplx *src_col = x.in + (size_t)row_l * col_stride;
      for (int c = 0; c < NCH; ++c, ++g) {
        const int s = g % NS;
        if (g >= NS) mbar_wait(&S.empty[s], ((g / NS) + 1) & 1);
        FwdStageWS<NTC, KC> &B = S.st[s];
        const int i0 = c * KC;
        const int cnt = min(KC, a.nrh - i0);         // radial points of this stage (table columns beyond are TMA zero-fill)
        unsigned bytes = row_ok ? (unsigned)cnt * 16u : 0u;
        if (tid == CONS) bytes += (unsigned)(KC / 16) * WSF_MT * 16u * 8u;
        mbar_arrive_expect_tx(&S.full[s], bytes);
        if (tid == CONS) {
#pragma unroll
          for (int bx = 0; bx < KC / 16; ++bx) tma_box3(&B.A[bx][0][0], &tmap, i0 + 16 * bx, x.n0, x.mglob, &S.full[s]);
        }
        if (row_ok && cnt < KC) {
          // radial points beyond nr/2 meet zero-filled table columns: keep stale NaN patterns out of the products
          double *row = is_top ? &B.T[row_l][2 * cnt] : &B.Bm[row_l][0];
          for (int q = 0; q < 2 * (KC - cnt); ++q) row[q] = 0.0;
        }
        if (row_ok) {
          if (is_top)
            bulk_g2s(&B.T[row_l][0], src_col + i0, (unsigned)cnt * 16u, &S.full[s]);
          else
            bulk_g2s(&B.Bm[row_l][2 * (KC - cnt)], src_col + (a.nr - i0 - cnt), (unsigned)cnt * 16u, &S.full[s]);
        }
      }
    }
  } else {
    // =============================== consumer warps ===============================
    // warp (h, wq): row tiles t = 2 mt + h (8 rows of each parity = 16 consecutive n), real columns 16 wq .. 16 wq + 15
    if (CREGS > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(CREGS));
    const int h = warp / NCG, wq = warp % NCG;
    const int fr = lane >> 2, fk = lane & 3;
    const int swap = a.swap_parity;   // 0: even rows contract with the sum fold (eomul), 1: with the difference (oemul)
    double acc[2][MT][2][2];

    // Columns without a retained row have no tile: they are zeros (se = 0, ops:1898-1899).  Written here, by the DMMA
    // warps, while the first stages are in flight.  (Rows beyond the last row tile of a column WITH retained rows
    // are zeroed by that tile's epilogue.)
    {
      const int nzero = a.npl - L.mcount;           // columns [0, mlo) and [mlo + mcount, npl)
      const int ncol = L.nfld * nzero * a.nzl;
      for (int col = blockIdx.x * Cfg::NCW + warp; col < ncol; col += gridDim.x * Cfg::NCW) {
        const int kz = col % a.nzl, rest = col / a.nzl;
        const int zc = rest % nzero, fld = rest / nzero;
        const int ml = zc < L.mlo ? zc : zc + L.mcount;
        cplx *o = a.fb.out[fld] + (size_t)kz * col_stride + (size_t)ml * a.nrl;
        for (int n = lane; n < a.nrdim; n += 32) o[n] = make_double2(0.0, 0.0);
      }
    }

    // NACT: active row tiles of this warp; SWAP: 0 = even rows contract with the sum fold (eomul), 1 = with the
    // difference (oemul); LN: remove the log term from the real part of the m = 0 column (ops:193-195)
    auto compute = [&](auto nact_c, auto swap_c, auto ln_c, const FwdStageWS<NTC, KC> &B, int i0, double lnval) {
      constexpr int NACT = decltype(nact_c)::value;
      constexpr int SWAP = decltype(swap_c)::value;
      constexpr bool LN = decltype(ln_c)::value != 0;
#pragma unroll
      for (int ks = 0; ks < KC / 4; ++ks) {
        const int k = ks * 4 + fk;                  // radial point within the stage
        const int kk = k & 15, box = k >> 4;
        double af[2][MT], bf[2][2];
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            if (mt < NACT) {
              const int nl = 16 * (HW * mt + h) + 2 * fr + p;
              af[p][mt] = B.A[box][nl][(((kk >> 1) ^ (nl & 7)) << 1) + (kk & 1)];
            }
        double l1 = 0.0, l2 = 0.0;
        if (LN) {
          const int i = i0 + k;
          if (i < a.nrh && !(fr & 1)) {             // real columns only
            l1 = lnval * __ldg(&a.lnx[i]);
            l2 = lnval * __ldg(&a.lnx[a.nr - 1 - i]);
          }
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int col = wq * 16 + nt * 8 + fr;    // real column: plane col / 2, real or imaginary part
          double t = B.T[col >> 1][2 * k + (col & 1)];
          double mr = B.Bm[col >> 1][2 * (KC - 1 - k) + (col & 1)];
          if (LN) {
            t -= l1;
            mr -= l2;
          }
          bf[SWAP][nt] = t + mr;
          bf[SWAP ^ 1][nt] = t - mr;
        }
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            if (mt < NACT) {
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) dmma884(acc[p][mt][nt][0], acc[p][mt][nt][1], af[p][mt], bf[p][nt]);
            }
      }
    };
    auto compute_n = [&](auto swap_c, auto ln_c, int nact, const FwdStageWS<NTC, KC> &B, int i0, double lnval) {
      if (MT == 4 && nact == 4) compute(IntK<(MT == 4 ? 4 : 1)>{}, swap_c, ln_c, B, i0, lnval);
      else if (MT == 4 && nact == 3) compute(IntK<(MT == 4 ? 3 : 1)>{}, swap_c, ln_c, B, i0, lnval);
      else if (nact == 2) compute(IntK<2>{}, swap_c, ln_c, B, i0, lnval);
      else if (nact == 1) compute(IntK<1>{}, swap_c, ln_c, B, i0, lnval);
    };

    int g = 0;
    for (int jj = 0;; ++jj) {
      const FwdCtxWS &x = S.ring[jj & (WS_RING - 1)];
      if (!x.valid) break;
      const int nn = x.nn, n0 = x.n0, kz0 = x.kz0;
      cplx *outp = x.out;
      const double lnval = x.lnval;
      const int nact = min(MT, max(0, (nn - n0 - 16 * h + 16 * HW - 1) / (16 * HW)));
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) acc[p][i][j][0] = acc[p][i][j][1] = 0.0;
      for (int c = 0; c < NCH; ++c, ++g) {
        const int s = g % NS;
        mbar_wait(&S.full[s], (g / NS) & 1);
        if (__double_as_longlong(lnval) != 0) {
          if (swap) compute_n(IntK<1>{}, IntK<1>{}, nact, S.st[s], c * KC, lnval);
          else compute_n(IntK<0>{}, IntK<1>{}, nact, S.st[s], c * KC, lnval);
        } else {
          if (swap) compute_n(IntK<1>{}, IntK<0>{}, nact, S.st[s], c * KC, lnval);
          else compute_n(IntK<0>{}, IntK<0>{}, nact, S.st[s], c * KC, lnval);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.empty[s]);
      }
      // thread holds C[row = tile*8 + lane/4][cols 2*(lane%4), +1] of each 8x8 tile == one complex per parity: rows n, n+1
      // of one z plane (32 contiguous bytes).  The table boxes carry real values beyond nn(m): those rows are zeros
      // of the truncated expansion.  (Swapping one value between neighbouring lanes so that every store instruction
      // writes whole 32-byte sectors changed nothing in isolation and cost 4 % inside the round trip: the stores are
      // not what the epilogue waits for.)
      {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int n = n0 + 2 * ((HW * mt + h) * 8 + fr);
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const int kz = kz0 + wq * 8 + nt * 4 + fk;
            if (kz < a.nzl) {
              cplx *o = outp + (size_t)kz * col_stride + n;
              if (n < a.nrdim)
                o[0] = (n < nn) ? make_double2(acc[0][mt][nt][0], acc[0][mt][nt][1]) : make_double2(0.0, 0.0);
              if (n + 1 < a.nrdim)
                o[1] = (n + 1 < nn) ? make_double2(acc[1][mt][nt][0], acc[1][mt][nt][1]) : make_double2(0.0, 0.0);
            }
          }
        }
      }
      // last row tile of its column: the rows beyond it are zeros (nr .. nrdim-1 when nrchop = nr)
      if (n0 + WSF_MT >= nn && n0 + WSF_MT < a.nrdim) {
        const int kz = kz0 + wq * 8 + (lane >> 2);
        if (kz < a.nzl) {
          cplx *o = outp + (size_t)kz * col_stride;
          for (int n = n0 + WSF_MT + 4 * h + (lane & 3); n < a.nrdim; n += 4 * HW) o[n] = make_double2(0.0, 0.0);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int g_ws_sms = 0;
static unsigned int *g_ws_ctr = nullptr;
static std::map<const double *, CUtensorMap> g_tmaps;   // one tensor map per analysis table of the current kit

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// big shape: 32 planes per tile, one CTA per SM; small shape: 16 planes per tile, two CTAs per SM
#define WSB_BIG 32, 6, 1
#define WSB_SMALL 16, 4, 2
// forward variants <NTC, KC, NS, MINB, HW, CREGS>; CREGS > 0: the DMMA warpgroups raise their register allowance to
// CREGS (setmaxnreg) once the producer warpgroup has dropped to 40.  The registers come out of the CTA's own allocation
// (threads x registers at launch): 640 x 96 here, so 104 is the most sixteen DMMA warps can get (112 never gets its
// registers and the CTA hangs).
#define WSF_BIG 32, 32, 3, 1, 2, 0        // 8 DMMA warps, 64 rows x 8 planes each
#define WSF_SMALL 16, 32, 2, 2, 2, 0      // two CTAs per SM
#define WSF_W16 32, 32, 3, 1, 4, 104      // 16 DMMA warps, 32 rows x 8 planes each: four per SM sub-partition

template <int NTC, int KC, int NS, int MINB, int HW, int CREGS>
static int fwd_set_smem() {
  CUDA_TRY(cudaFuncSetAttribute(leg_forward_ws_kernel<NTC, KC, NS, MINB, HW, CREGS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(FwdSmemWS<NTC, KC, NS>)));
  return MLEGS_OK;
}
template <int NTC, int KC, int NS, int MINB, int HW, int CREGS>
static void fwd_launch(int grid, const LegArgs &b, const LegItemsF &L, const CUtensorMap &tm, cudaStream_t st) {
  leg_forward_ws_kernel<NTC, KC, NS, MINB, HW, CREGS><<<grid, WsCfg<NTC, HW>::THREADS, sizeof(FwdSmemWS<NTC, KC, NS>), st>>>(b, L, tm);
}

static int ws_setup() {
  if (g_ws_sms) return MLEGS_OK;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&g_ws_sms, cudaDevAttrMultiProcessorCount, dev));
  CUDA_TRY(cudaMalloc((void **)&g_ws_ctr, sizeof(unsigned int)));
  const int big_b = (int)sizeof(BwdSmemWS<32, 6>), small_b = (int)sizeof(BwdSmemWS<16, 4>);
  CUDA_TRY(cudaFuncSetAttribute(leg_backward_ws_kernel<0, WSB_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_b));
  CUDA_TRY(cudaFuncSetAttribute(leg_backward_ws_kernel<1, WSB_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_b));
  CUDA_TRY(cudaFuncSetAttribute(leg_backward_ws_kernel<2, WSB_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_b));
  CUDA_TRY(cudaFuncSetAttribute(leg_backward_ws_kernel<0, WSB_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_b));
  CUDA_TRY(cudaFuncSetAttribute(leg_backward_ws_kernel<1, WSB_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_b));
  CUDA_TRY(cudaFuncSetAttribute(leg_backward_ws_kernel<2, WSB_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_b));
  MLEGS_TRY(fwd_set_smem<WSF_BIG>());
  MLEGS_TRY(fwd_set_smem<WSF_SMALL>());
  MLEGS_TRY(fwd_set_smem<WSF_W16>());
  return MLEGS_OK;
}

// Tile shape of a launch; MLEGS_LEG_SHAPE=big|small forces one for A/B timing.
static bool ws_small_shape(const LegArgs &a, bool forward) {
  static const char *force = getenv("MLEGS_LEG_SHAPE");
  if (force) return force[0] == 's';
  // measured (tools/leg_bench.py, profiles/r2): synthesis 128^3 75.8 us small / 77.1 big, 256^3 905 / 876; analysis
  // 128^3 92.5 / 87.2 (the small shape re-reads the table slice twice as often and the analysis is bound by its
  // L2 -> shared-memory fills), 256^3 1023 / 989
  return !forward && a.nrh <= 64;
}

// table(i, n, m), i fastest: boxes of 16 i x 128 n x 1 m, 128-byte swizzle, zero fill outside the table
static int table_tmap(const double *tab, int nrh, int ne, int nm, const CUtensorMap **out) {
  auto it = g_tmaps.find(tab);
  if (it == g_tmaps.end()) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
      void *fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      if (!fn || qres != cudaDriverEntryPointSuccess) return fail(MLEGS_E_CUDA, "cuTensorMapEncodeTiled is not available");
      encode = (EncodeTiledFn)fn;
    }
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)nrh, (cuuint64_t)ne, (cuuint64_t)nm};
    const cuuint64_t strides[2] = {(cuuint64_t)nrh * 8, (cuuint64_t)nrh * ne * 8};
    const cuuint32_t box[3] = {16, WSF_MT, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)tab, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MLEGS_E_CUDA, "cuTensorMapEncodeTiled failed for the Legendre table");
    it = g_tmaps.emplace(tab, tm).first;
  }
  *out = &it->second;
  return MLEGS_OK;
}

__global__ void pfw_kernel(const double *__restrict__ pf, const double *__restrict__ w, double *__restrict__ out, int nrh,
                           size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = pf[i] * w[i % nrh];
}

// pf * w, built once per kit (the reference multiplies the weights into the data of every transform, ops:1903-1907)
int leg_ws_build_pfw() {
  Context &c = ctx();
  const size_t n = (size_t)c.nrh * c.ne * c.p.npchop;
  CUDA_TRY(cudaMalloc((void **)&c.d_pfw, n * sizeof(double)));
  pfw_kernel<<<1184, 256, 0, (cudaStream_t)c.stream>>>(c.d_pf, c.d_w, c.d_pfw, c.nrh, n);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)c.stream));
  return MLEGS_OK;
}

void leg_ws_reset() { g_tmaps.clear(); }

// the bulk copies need 16-byte aligned table rows: nr a multiple of 4
bool leg_ws_supported(const LegArgs &a) { return (a.nrh & 1) == 0; }

int launch_slab_ship(const PeerTable &t, const FieldBatch &fb, cudaStream_t st);   // dist.cu

// Algorithmic work of one radial transform launch: flops 2 nr nz S with S = sum of nn(m) over the
// local columns (parity-folded real table x complex data); bytes 16 nr nz per column with work + 16 S nz.
void leg_alg_work(const LegArgs &a, double *bytes, double *flops) {
  double S = 0.0, cols = 0.0;
  for (int ml = 0; ml < a.npl; ++ml) {
    const int m = a.m0 + ml * a.ms;
    const int nn = m < a.npc ? std::max(std::min(a.nrc, a.nrc - m), 0) : 0;
    if (nn > 0 && !(a.skip_m0 && m == 0)) {
      S += nn;
      cols += 1.0;
    }
  }
  const double nf = a.fb.n > 0 ? a.fb.n : 1;
  *flops = nf * 2.0 * a.nr * a.nzl * S;
  *bytes = nf * (16.0 * a.nr * cols * a.nzl + 16.0 * S * a.nzl);
}

template <int NTC, int NS, int MINB>
static void launch_bwd_shape(int mode, int grid, const LegArgs &a, const LegItemsB &L, const PeerTable &pt, cudaStream_t st) {
  const size_t smem = sizeof(BwdSmemWS<NTC, NS>);
  const int threads = WsCfg<NTC>::THREADS;
  if (mode == 2)
    leg_backward_ws_kernel<2, NTC, NS, MINB><<<grid, threads, smem, st>>>(a, L, pt);
  else if (mode == 1)
    leg_backward_ws_kernel<1, NTC, NS, MINB><<<grid, threads, smem, st>>>(a, L, pt);
  else
    leg_backward_ws_kernel<0, NTC, NS, MINB><<<grid, threads, smem, st>>>(a, L, pt);
}

// `a` has its batch filled in (legendre.cu: with_batch)
int launch_leg_backward_ws(const LegArgs &a, cudaStream_t st) {
  MLEGS_TRY(ws_setup());
  const bool small = ws_small_shape(a, false);
  const int ntc = small ? 16 : 32;
  LegItemsB L;
  L.nkz = (a.nzl + ntc - 1) / ntc;
  L.nit = (a.nrh + WSB_MT - 1) / WSB_MT;
  L.nfld = a.fb.n;
  L.mcount = 0;
  for (int ml = 0; ml < a.npl; ++ml) {   // nn(m) is non-increasing in m: the columns with work form a prefix
    const int m = a.m0 + ml * a.ms;
    const int nn = m < a.npc ? std::max(std::min(a.nrc, a.nrc - m), 0) : 0;
    if (nn <= 0) break;
    ++L.mcount;
  }
  L.total = L.mcount * L.nfld * L.nit * L.nkz;
  // even without a tile the grid covers the zero-filled columns
  const int grid = std::max(1, std::min(g_ws_sms * (small ? 2 : 1), std::max(L.total, a.npl > L.mcount ? g_ws_sms : 1)));
  // several ranks: with an output buffer per scalar the rows are staged locally in destination order and shipped as
  // long runs (launch_slab_ship); without one they are put straight into the peers' windows
  // (2 ranks, 128^3 per rank: direct puts 192 us, staged 120 + 133 us; 8 ranks: the direct puts are 256-byte runs of
  // 128-byte stores and reach 243 GB/s).  Staged when a peer's share of the rows is shorter than 48 (768-byte runs).
  static const char *force = getenv("MLEGS_EXCHANGE12");   // "stage" / "put": A/B timing
  bool stage = a.peer && a.fb.out[0] != nullptr && (a.nrdim / a.peer->nranks) < 48;
  if (a.peer && a.fb.out[0] != nullptr && force) stage = force[0] == 's';
  double wb = 0.0, wf = 0.0;
  leg_alg_work(a, &wb, &wf);
  prof_begin(a.peer ? (stage ? "legendre_backward_stage" : "legendre_backward_put") : "legendre_backward", st, wb, wf);
  PeerTable none;
  memset(&none, 0, sizeof(none));
  const PeerTable &pt = a.peer ? *a.peer : none;
  const int mode = stage ? 2 : (a.peer ? 1 : 0);
  if (small)
    launch_bwd_shape<WSB_SMALL>(mode, grid, a, L, pt, st);
  else
    launch_bwd_shape<WSB_BIG>(mode, grid, a, L, pt, st);
  prof_end(st);
  KERNEL_CHECK();
  if (stage) MLEGS_TRY(launch_slab_ship(*a.peer, a.fb, st));
  return MLEGS_OK;
}

// forward: the table contracted against is a.pf (w == nullptr: the vec2tp projection tables) or pf*w
bool leg_ws_forward_supported(const LegArgs &a) {
  Context &c = ctx();
  if ((a.nrh & 1) != 0) return false;
  if (a.w != nullptr && !(a.pf == c.d_pf && a.w == c.d_w && c.d_pfw)) return false;
  return true;
}

int launch_leg_forward_ws(const LegArgs &a, cudaStream_t st) {
  Context &c = ctx();
  MLEGS_TRY(ws_setup());
  const double *tab = a.w ? c.d_pfw : a.pf;
  const CUtensorMap *tm = nullptr;
  MLEGS_TRY(table_tmap(tab, a.nrh, a.ne, c.p.npchop, &tm));
  const bool small = ws_small_shape(a, true);
  const int ntc = small ? 16 : 32;
  // Sixteen DMMA warps for short contractions (tools/leg_bench.py, 8 scalars per launch: 128^3 80.0-81.9 us against
  // 83.9-85.1 with eight; 256^3 963 / 979; 512^3 13.91 / 13.82 ms: the epilogue and tile turn-over that more warps
  // shorten weigh less the longer a tile is); MLEGS_LEG_FWD=8|16 forces one for A/B timing
  static const char *fv = getenv("MLEGS_LEG_FWD");
  const bool w16 = !small && (fv ? atoi(fv) == 16 : a.nrh <= 128);
  const int kc = 32;
  // tiles: (column with nn(m) > 0) x scalar x z tile x row tile of the widest column; nn(m) is non-increasing in m,
  // so the columns with work form a prefix
  LegItemsF L;
  L.nkz = (a.nzl + ntc - 1) / ntc;
  L.nfld = a.fb.n;
  L.mlo = (a.skip_m0 && a.m0 == 0) ? 1 : 0;
  auto nn_host = [&](int ml) {
    const int m = a.m0 + ml * a.ms;
    return m < a.npc ? std::min(std::max(std::min(a.nrc, a.nrc - m), 0), a.nrdim) : 0;
  };
  int mcount = 0;
  for (int ml = L.mlo; ml < a.npl && nn_host(ml) > 0; ++ml) ++mcount;
  L.nrt = mcount > 0 ? (nn_host(L.mlo) + WSF_MT - 1) / WSF_MT : 1;
  const int total = mcount * L.nfld * L.nkz * L.nrt;
  L.total = total;
  L.mcount = mcount;
  L.ctr = nullptr;
  if (!small && (a.nrh + kc - 1) / kc >= 8) {
    L.ctr = g_ws_ctr;
    CUDA_TRY(cudaMemsetAsync(g_ws_ctr, 0, sizeof(unsigned int), st));
  }
  // even without a tile the grid covers the zero-filled columns
  const int grid = std::max(1, std::min(g_ws_sms * (small ? 2 : 1), std::max(total, a.npl > mcount ? g_ws_sms : 1)));
  LegArgs b = a;
  b.pf = tab;
  double wb = 0.0, wf = 0.0;
  leg_alg_work(a, &wb, &wf);
  prof_begin("legendre_forward", st, wb, wf);
  if (small)
    fwd_launch<WSF_SMALL>(grid, b, L, *tm, st);
  else if (w16)
    fwd_launch<WSF_W16>(grid, b, L, *tm, st);
  else
    fwd_launch<WSF_BIG>(grid, b, L, *tm, st);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

}  // namespace mlegs
