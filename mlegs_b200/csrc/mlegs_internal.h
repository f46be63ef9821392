// Internal declarations shared by the translation units of libmlegs_b200.so.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mlegs_b200.h"

#ifdef __CUDACC__
#include <cuda_runtime.h>
#endif

namespace mlegs {

// ---- error handling ------------------------------------------------------------------
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define MLEGS_TRY(expr)                \
  do {                                 \
    int _rc = (expr);                  \
    if (_rc != MLEGS_OK) return _rc;   \
  } while (0)

#ifdef __CUDACC__
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
#define CUDA_TRY(expr)                                                       \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) return ::mlegs::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)
#define KERNEL_CHECK()                                                       \
  do {                                                                       \
    ::mlegs::g_launches++;                                                   \
    cudaError_t _e = cudaGetLastError();                                     \
    if (_e != cudaSuccess) return ::mlegs::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)
#endif

extern long long g_launches;

// ---- host table builder (tfm_tables.cpp) --------------------------------------------------
int build_tfm_tables(const mlegs_params *p, double *x, double *w, double *ln, double *r, double *lognorm,
                     double *pf, double *at0, double *at1, double *ak);

int build_tfm_tables_cached(const mlegs_params *p, const char *cache_dir, double *x, double *w, double *ln, double *r,
                            double *lognorm, double *pf, double *at0, double *at1, double *ak, int *from_cache);

// ---- FFT plan: radix schedule + twiddles for one length ------------------------------------
struct FftPlan {
  int n = 0;            // complex length of the in-smem FFT
  int npass = 0;
  int radix[16] = {0};
  int ti = 0;           // lines per CTA
  size_t smem = 0;      // dynamic shared memory bytes
};

// ---- the context: everything tfm%init() + module globals hold, resident on the device ---------
struct Context {
  bool ready = false;
  mlegs_params p{};
  int rank = 0, nranks = 1;
  int nrdim = 0, npdim = 0, nzdim = 0;
  int nrh = 0, ne = 0;          // nr/2, nrchop+14
  int chopzl = 0, chopzu = 0;
  // host copies
  std::vector<double> h_x, h_w, h_ln, h_lognorm, h_at0, h_at1, h_ak;
  std::vector<double> h_del2h, h_xxdx;   // band tables, see d_del2h / d_xxdx
  int *d_flag = nullptr;                 // device error/finite flags [0]: non-finite, [1]: singular pivot
  // device tables
  double *d_x = nullptr, *d_w = nullptr, *d_lnx = nullptr, *d_r = nullptr;
  double *d_lognorm = nullptr;   // (ne, npchop)
  double *d_pf = nullptr;        // (nrh, ne, npchop)
  double *d_pfw = nullptr;       // pf(i,n,m) * w(i): the analysis table of the TMA-fed Legendre kernel
  double *d_at0 = nullptr, *d_at1 = nullptr, *d_ak = nullptr;
  double *d_tw_p = nullptr;      // np twiddles  exp(-2 pi i j/np), interleaved re,im
  double *d_tw_z = nullptr;      // nz twiddles
  // band tables, per m: coefficient d of row n is at [(d_idx * ne + n) + m * nb * ne]
  double *d_del2h = nullptr;     // 5 diagonals (-2..2), lognorm-scaled, /ell^2   (sdiff:45-96)
  double *d_xxdx = nullptr;      // 3 diagonals (-1..1)                           (sdiff:6-43)
  // projection tables for vec2tp (ops:1405-1411), same shape as pf but nrchop+3 columns used
  double *d_vtab = nullptr, *d_dtab = nullptr;
  // scratch
  void *d_scratch[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // field-sized work buffers
  void *d_batch[32] = {};   // trans_many, lazy
  double *d_cossin_p = nullptr;  // cos / sin of the np azimuthal collocation angles (on-device initial conditions)
  static const int RED_DOUBLES = 32768;
  size_t field_bytes = 0;
  double *d_red = nullptr;       // small reduction workspace
  double *h_red = nullptr;       // pinned
  void *d_solve_ws = nullptr;    // banded-solve workspace (U factors)
  size_t solve_ws_bytes = 0;
  FftPlan plan_p, plan_z;
  // compact axial FFT: prefix sums of nn(m) over the local columns, cached per (nrc, npc)
  int *d_colstart = nullptr;
  int cs_nrc = -1, cs_npc = -1, cs_ncols = 0;
  long long cs_total = 0;
  void *stream = nullptr;        // cudaStream_t
  // multi-GPU
  void *d_window = nullptr;      // exchange window (field-sized), exported over CUDA IPC
  std::vector<void *> peer_window;
  std::vector<void *> peer_flags;
  void *d_flags = nullptr;
  unsigned long long epoch = 0;
  // r: decompose() shares of nrdim (mlegs_envir_mpi.f90:6-31).  m: CYCLIC ownership -- rank q holds the columns
  // m = q, q + P, ... (m_cnt[q] of them) -- because the work per column falls linearly with m (nn(m) = nrchop - m):
  // contiguous blocks give rank 0 1.4x the mean radial-transform / axial-FFT / solve work at 8 ranks.  m_off = prefix
  // sums of m_cnt: where rank q's columns sit in an exchange window that groups columns by source rank.
  std::vector<int> r_cnt, r_off, m_cnt, m_off;
};

Context &ctx();

// decompose, submodules/mlegs_envir_mpi.f90:6-31
inline void decompose(int nsize, int nprocs, int proc, int *cnt, int *off) {
  int q = nsize / nprocs, r = nsize % nprocs;
  if (r > proc) {
    *cnt = q + 1;
    *off = (q + 1) * proc;
  } else {
    *cnt = q;
    *off = q * proc + r;
  }
}

// chop_index, ops:2012-2063
struct ChopIdx {
  int nrc, npc, nzc, nzcu;
};
int chop_index(const mlegs_field *s, ChopIdx *ci);
void field_set_layout(mlegs_field *f, bool physical);
// global m of local column j of a block = loc_st[1] + j * field_mstride: 1 when the block holds every azimuthal column,
// the number of ranks when m is distributed (cyclic ownership: rank q holds m = q, q + P, q + 2P, ...)
int field_mstride(const mlegs_field *f);
int validate_params(const mlegs_params *p);

}  // namespace mlegs
