// Kernel launchers (device code lives in fft.cu, legendre.cu, elementwise.cu, banded.cu).
#pragma once
#include <cuda_runtime.h>

#include "mlegs_internal.h"

namespace mlegs {

typedef double2 cplx;

// prof.cu: optional CUDA-event timing around each launch
void prof_begin(const char *name, cudaStream_t st);
void prof_end(cudaStream_t st);

// ---- fft.cu ----------------------------------------------------------------------------
int make_fft_plan(int n_complex, int extra_points, FftPlan *plan);
int setup_fft_kernels();
// batched strided line FFTs.  A "line" has points at element stride `stride_pt`; lines are
// batched contiguously (`batch0` consecutive elements) and then by `batch1` blocks at stride_b1.
enum FftMode { FFT_C2C_FWD = 0, FFT_C2C_BWD = 1, FFT_R2C_FWD = 2, FFT_C2R_BWD = 3 };
int launch_fft_lines(FftMode mode, const FftPlan &plan, const cplx *in, cplx *out, long long batch0,
                     long long stride_pt, int batch1, long long stride_b1, const double *tw, int tw_order,
                     double scale, cudaStream_t st);

// ---- legendre.cu -----------------------------------------------------------------------
struct LegArgs {
  const cplx *in;
  cplx *out;
  const double *pf;      // (nrh, ne, npchop)
  const double *w;       // nr
  const double *lnx;     // nr, -log(1-x)
  int nr, nrh, ne;
  int nrl;               // leading dimension (rows) of in/out == nrdim
  int npl;               // local number of m columns
  int m0;                // global m of local column 0
  int nzl;               // number of z planes (complex columns per m)
  int nrc, npc;          // chop limits incl. offsets: nn(m) = max(min(nrc, nrc-m),0) for m < npc
  int nrdim;
  double lnval;          // s%ln (log-term), applied on global m == 0
};
int setup_leg_kernels();
int launch_leg_forward(const LegArgs &a, cudaStream_t st);
int launch_leg_backward(const LegArgs &a, cudaStream_t st);

}  // namespace mlegs
