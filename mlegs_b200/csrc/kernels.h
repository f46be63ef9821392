// Kernel launchers (device code lives in fft.cu, legendre.cu, elementwise.cu, banded.cu).
#pragma once
#include <cuda_runtime.h>

#include "mlegs_internal.h"

namespace mlegs {

typedef double2 cplx;

// Several independent scalars of identical layout handled by ONE launch (the three components of a vector field,
// ops:1503-1505; a batch of scalars in mlegs_b200_trans_many): the field index is a grid dimension.
#define MLEGS_MAXB 32
struct FieldBatch {
  int n = 0;
  const cplx *in[MLEGS_MAXB];
  cplx *out[MLEGS_MAXB];
  double ln[MLEGS_MAXB];
};

// Row scaling fused into the azimuthal FFT of some scalars of a batch: rows r0 + i < nr are multiplied by r(i) as they
// are loaded (mode 1: the r*u of vec2tp, ops:1337-1355, on the r2c side) or divided by r(i) as they are stored
// (mode 2: the u/r of tp2vec, ops:1509-1527, on the c2r side).  Same operations in the same order as a separate
// rscale pass, so results are bit-identical.
struct RowScale {
  const double *r = nullptr;
  int mode = 0;
  unsigned mask = 0;     // bit i: scalar i of the batch is scaled
  int r0 = 0, nr = 0;
  // c2r side only: the input is an exchange window in the transit layout of the fused exchange(1,2) -- columns grouped
  // by source rank (dist_dev.cuh) -- so point m of a line is read at column perm_off[m mod perm_p] + m / perm_p
  int perm_p = 0;
  int perm_off[16];
};

// prof.cu: optional CUDA-event timing around each launch
// bytes / flops: the ALGORITHMIC HBM bytes and floating-point operations of this launch, summed
// per kernel name in the report so that bench.py can quote achieved GB/s and TFLOP/s for every kernel class
void prof_begin(const char *name, cudaStream_t st, double bytes = 0.0, double flops = 0.0);
void prof_end(cudaStream_t st);

// elements (n, m, k) of a local (nrl, npl, nzl) spectral block that the truncation keeps: rows r0 + i < nn(m) of the
// columns m0 + j < npc, planes k < nzc or k >= nzcu (host helper of the algorithmic-byte counts above)
inline double retained_elems(int nrl, int npl, int nzl, int r0, int m0, int nrc, int npc, int nzc, int nzcu, int ms = 1) {
  double rows = 0.0;
  for (int j = 0; j < npl; ++j) {
    const int m = m0 + j * ms;
    int nn = m < npc ? (nrc < nrc - m ? nrc : nrc - m) : 0;
    nn -= r0;
    rows += nn < 0 ? 0 : (nn > nrl ? nrl : nn);
  }
  int planes = 0;
  for (int k = 0; k < nzl; ++k) planes += (k < nzc || k >= nzcu) ? 1 : 0;
  return rows * planes;
}

// ---- fft.cu ----------------------------------------------------------------------------
int make_fft_plan(int n_complex, int extra_points, FftPlan *plan);
int setup_fft_kernels();
// batched strided line FFTs.  A "line" has points at element stride `stride_pt`; lines are
// batched contiguously (`batch0` consecutive elements) and then by `batch1` blocks at stride_b1.
enum FftMode { FFT_C2C_FWD = 0, FFT_C2C_BWD = 1, FFT_R2C_FWD = 2, FFT_C2R_BWD = 3 };
int launch_fft_lines(FftMode mode, const FftPlan &plan, const cplx *in, cplx *out, long long batch0,
                     long long stride_pt, int batch1, long long stride_b1, const double *tw, int tw_order,
                     double scale, cudaStream_t st, const FieldBatch *fb = nullptr, const RowScale *rs = nullptr);

// fft_reg.cu: register-resident fast path for power-of-two lengths 32..1024
bool fft_reg_supported(int n);
struct PeerTable;   // dist_dev.cuh
int launch_fft_reg(FftMode mode, int n, const cplx *in, cplx *out, long long nlines, long long batch0,
                   long long stride_b1, long long stride_pt, const double *tw, int tw_order, double scale,
                   const int *colstart, int ncols, int nrl, cudaStream_t st, const PeerTable *peer = nullptr,
                   int nrdim = 0, const FieldBatch *fb = nullptr, const RowScale *rs = nullptr);
// azimuthal r2c FFT whose stores are the exchange(2,1) puts (several ranks, register kernels only)
int launch_fft_phi_forward_put(const FftPlan &plan, const cplx *in, long long rows, int nz, long long plane,
                               const double *tw, int tw_order, double scale, const PeerTable &peer, int nrdim,
                               cudaStream_t st, const FieldBatch *fb = nullptr, const RowScale *rs = nullptr);
// axial FFT of the retained lines only (rows < nn(m) of each local column); colstart = device prefix sums
int launch_fft_z_compact(FftMode mode, const FftPlan &plan, const cplx *in, cplx *out, const int *colstart, int ncols,
                         int nrl, long long nlines, long long stride_pt, const double *tw, int tw_order, double scale,
                         cudaStream_t st, const FieldBatch *fb = nullptr);

// ---- legendre.cu -----------------------------------------------------------------------
struct LegArgs {
  const cplx *in;
  cplx *out;
  const double *pf;      // (nrh, ne, npchop) table contracted against (pf, or the v/d projection tables)
  const double *w;       // nr quadrature weights (forward only; nullptr: no weighting)
  const double *lnx;     // nr, -log(1-x)
  int nr, nrh, ne;
  int nrl;               // leading dimension (rows) of in/out == nrdim
  int npl;               // local number of m columns
  int m0;                // global m of local column 0
  int ms = 1;            // global m of local column j = m0 + j ms (ms = number of ranks when m is distributed: cyclic)
  int nzl;               // number of z planes (complex columns per m)
  int nrc, npc;          // chop limits incl. offsets: nn(m) = max(min(nrc, nrc-m),0) for m < npc
  int nrdim;
  double lnval;          // s%ln (log-term), applied on global m == 0
  int swap_parity;       // forward only: 0 = eomul (even rows <- even fold), 1 = oemul (ops:2067-2145)
  int skip_m0;           // forward only: leave the m == 0 column zero (vec2tp: `if (mv .ne. 0)`)
  const PeerTable *peer; // backward only (host pointer, nullptr on one rank): the stores are the exchange(1,2) puts
  int npdim;             // global number of m columns (fused put addressing)
  FieldBatch fb;         // fb.n > 0: the launch handles fb.n scalars (in/out/lnval above are ignored)
};
int setup_leg_kernels();
int launch_leg_forward(const LegArgs &a, cudaStream_t st);
int launch_leg_backward(const LegArgs &a, cudaStream_t st);

// ---- elementwise.cu ---------------------------------------------------------------------
struct MaskArgs {
  cplx *e;
  int nrl, npl, nzl, r0, m0;
  int ms = 1;            // column j holds m = m0 + j ms
  int row_mode;
  int nrc, npc_rows;
  int col_cut;
  int kz_lo, kz_hi;
};
int launch_mask(const MaskArgs &a, cudaStream_t st);

struct SvvArgs {
  cplx *e;
  int nrl, npl, nzl, r0, m0;
  int ms = 1;
  const double *ak;
  int nak;
  double qr_den, qp_den, kmax, cutoff, strength;
};
int launch_svv_energy(const SvvArgs &a, double *d_partial, double *d_out2, cudaStream_t st);
int launch_svv_apply(const SvvArgs &a, cudaStream_t st);

int launch_calcat(cplx *e, int nrl, int npl, int nzl, int nrows, const double *at, cplx *out, int subtract,
                  double at_first, cudaStream_t st);
int launch_delsqp(cplx *e, int nrl, int npl, int nzl, int m0, int ms, int nrc, int npc, double ell2, int inverse,
                  cudaStream_t st);
struct PokeArgs {
  int n;
  long long off[4];
  double re[4], im[4];
  int mode[4];   // 0: set, 1: add
};
int launch_poke(cplx *e, const PokeArgs &p, cudaStream_t st);
int launch_zero_line(cplx *e, long long off, long long stride, int n, cudaStream_t st);
int launch_vecprod(cplx *vr, cplx *vp, cplx *vz, const cplx *ur, const cplx *up, const cplx *uz, int nrl, int npl,
                   int nzl, int r0, int nr, int nph, int nz, cudaStream_t st);
struct LinArgs {
  int mode;
  size_t n;
  cplx *y;
  const cplx *x1, *x2, *x3, *x4 = nullptr;
  double a, b, c, d = 0.0;
};
int launch_lincomb(const LinArgs &p, cudaStream_t st);
int launch_rscale(cplx *e, int nrl, size_t ncols, int r0, int nr, const double *r, int divide, cudaStream_t st);
int launch_finite(const cplx *e, size_t n, int *flag, cudaStream_t st);
int launch_col_update(cplx *col, int n, int mode, const double *v1, const double *v2, double s, cudaStream_t st);

// fftreat's far-field treatment of the radially synthesised ('PFF') array, ops:1023-1054, one CTA per (m,k) line
struct FftreatArgs {
  cplx *e;
  int nrl, npl, nzl, m0;
  int ms = 1;
  int nr, ns, ns0;                       // 1-based ns = nr*3/4, ns0 = min(ns+4, nr) of ops:1015-1016
  int npc, nzc, nzcu;
  const double *x;                       // Gauss-Legendre nodes
};
int launch_fftreat_tail(const FftreatArgs &a, cudaStream_t st);

// vec2tp combination (ops:1413-1435) and tp2vec combination (ops:1488-1502)
struct TpCombineArgs {
  cplx *dst;                             // psi or chi, rows < nn of the retained (m,k) columns only
  const cplx *t;                         // one of eomul(v,ur), oemul(d,up), oemul(d,ur), eomul(v,up), eomul(t,uz)
  int mode;                              // 0: dst=(-iu*mv)*t  1: dst-=t  2: dst=(iu*kv)*t  3: dst+=(mv*kv)*t  4: dst-=t
  int nrl, npl, nzl, m0;
  int ms = 1;
  int nrc, npc, nzc, nzcu;
  const double *ak;
};
int launch_tp_combine(const TpCombineArgs &a, cudaStream_t st);
struct TvCombineArgs {
  cplx *ur, *up;                         // in: xxdx(chi), xxdx(psi); out: combined
  const cplx *psi, *uz;                  // psi, chi
  int nrl, npl, nzl, m0;
  int ms = 1;
  int nrc, npc, nzc, nzcu;
  const double *ak;
};
int launch_tv_combine(const TvCombineArgs &a, cudaStream_t st);

// ---- banded.cu --------------------------------------------------------------------------
struct BandOpArgs {
  cplx *e;
  const cplx *src = nullptr;   // != nullptr: planes are read from here and EVERY line of the block is written to e
                               // (operator image where the operator acts, a copy elsewhere): replaces a field copy
  int out_neg = 0;             // the whole result is negated on the way out (the `s%e = -s%e` after del2, ops:1489, 1554)
  int nrl, npl, nzl, m0;
  int ms = 1;            // column j holds m = m0 + j ms
  const double *tab;     // (ne, nb, npchop) band coefficients
  int nb, ne;
  const double *ak;      // nullptr: no -ak^2 on the diagonal (xxdx, del2h)
  int nrc, npc, nzc, nzcu;
  int napply;            // 1, or power/2 for helmp
  int combine;           // helmp: out = (sp + beta*s2) + alpha*s on the WHOLE array (ops:893)
  double alpha, beta;
  int nlnc;              // log-term corrections added to rows 0..nlnc-1 of column (m=0,k=0)
  double lnc[3];
};
int launch_band_op(const BandOpArgs &a, cudaStream_t st);

struct SolveArgs {
  cplx *e;
  int nrl, npl, m0;
  int ms;                // column j holds m = m0 + j ms
  int k0, nk;            // axial planes k0 .. k0+nk-1
  const double *tab;     // del2h table
  int ne;
  const double *ak;
  int nrc, npc;
  int nnmax;
  int kl, ku;
  int power;             // 2: del2 (+alpha); 4,6,8: del^p + beta del2 + alpha
  int add_alpha;
  double alpha, beta;
  int special00;         // 0 none; 1 idel2_proln column fix (ops:705-711); 2 idel2_preln (ops:609-622)
  double sp0, sp1, sp2, preln_rhs;
  size_t ws_doubles;
  double *ws_global;
  int *flag;
  // factor cache (filled on the first solve with a given operator, reused by every later right-hand side)
  double *fac_ab;                // LU factors of all systems: the U rows (kl + ku + 1 per matrix column) of every column
                                 // first, then the L multipliers (kl per column) -- each substitution sweep streams
                                 // only its own part, contiguously (LAPACK's interleaved band storage made both sweeps
                                 // fetch 1.5x the factor bytes from HBM: 32-byte sectors of 200-byte columns)
  long long fac_ncols;           // matrix columns in the set: the L part starts at fac_ab + fac_ncols (kl + ku + 1)
  unsigned char *fac_piv;        // pivot offsets jp (0..kl) per column
  const long long *fac_off;      // per local column j: offset (in columns) of system (j, k0); see launch_band_solve
  // The operator of plane k only contains ak(k)^2 = ak(nz-k)^2: the cached substitution can serve plane nz-k with the
  // factors of plane k.  mirror_mode 1: every warp solves (j,k) and then (j, nz-k) if mirror_lo <= nz-k < mirror_nz;
  // 2: the mirrored planes only (launch_band_solve_ranges decides, banded.cu).
  int mirror_mode, mirror_lo, mirror_nz;
};
int launch_band_solve(SolveArgs a, cudaStream_t st);
// planes [0, n1) and [lo, nzl) of one operator (the two axial loops of the reference's solves)
int launch_band_solve_ranges(SolveArgs base, int n1, int kl_first, int lo, int nzl, cudaStream_t st);
void band_solve_cache_clear();
// 1: solves reuse cached LU factors (default); 0: every call factors again (the reference's behaviour)
void band_solve_cache_enable(int on);

}  // namespace mlegs
