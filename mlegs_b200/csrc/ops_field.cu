// Host-side bodies of the scalar operators declared in modules/mlegs_scalar.f90: masks and filters,
// spectral differential operators, Helmholtz/Poisson solves, time integrators and the toroidal-poloidal
// vector operations (/root/reference/src/submodules/mlegs_scalar_ops.f90:6-1560).  Each entry mirrors the
// reference procedure step for step and launches the kernels of elementwise.cu / banded.cu /
// legendre.cu / fft.cu; no arithmetic on field data happens on the host.
#include <climits>
#include <cmath>
#include <cstring>

#include "kernels.h"

namespace mlegs {

int trans_impl(mlegs_field *s, const char *to);
int dist_allreduce(double *d_inout, int n);   // dist.cu: sum over ranks, identical result everywhere
int dist_check_timeout();
int trans_many_impl(int n, mlegs_field *const *s, const char *to);
int stage_z(const mlegs_field *s, bool forward, const cplx *src, cplx *dst, const FieldBatch *fb = nullptr);
struct PeerTable;
int stage_r(const mlegs_field *s, bool forward, const cplx *src, cplx *dst, const PeerTable *peer = nullptr,
            const FieldBatch *fb = nullptr);

static cudaStream_t strm() { return (cudaStream_t)ctx().stream; }
static size_t nelem(const mlegs_field *f) { return (size_t)f->loc_sz[0] * f->loc_sz[1] * f->loc_sz[2]; }
static bool is_space(const mlegs_field *f, const char *sp) { return strncmp(f->space, sp, 3) == 0; }
static void set_space3(mlegs_field *f, const char *sp) {
  memcpy(f->space, sp, 3);
  f->space[3] = 0;
}
static int ready() {
  if (!ctx().ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  return MLEGS_OK;
}
static int require_fff(mlegs_field *s) {   // the reference warns and transforms (e.g. ops:338-342)
  if (!is_space(s, "FFF")) return trans_impl(s, "FFF");
  return MLEGS_OK;
}
static bool owns_m0(const mlegs_field *s) { return s->loc_st[1] == 0 && s->loc_sz[1] > 0; }

// a temporary scalar living in one of the context's scratch buffers, same layout/metadata as `like`
static mlegs_field temp_like(const mlegs_field *like, int slot) {
  mlegs_field t = *like;
  t.e = ctx().d_scratch[slot];
  return t;
}
static int copy_data(mlegs_field *dst, const mlegs_field *src) {
  void *e = dst->e;
  *dst = *src;
  dst->e = e;
  CUDA_TRY(cudaMemcpyAsync(dst->e, src->e, nelem(src) * sizeof(cplx), cudaMemcpyDeviceToDevice, strm()));
  return MLEGS_OK;
}
static int read_elem(const mlegs_field *s, size_t off, cplx *out) {
  CUDA_TRY(cudaMemcpyAsync(out, (const cplx *)s->e + off, sizeof(cplx), cudaMemcpyDeviceToHost, strm()));
  CUDA_TRY(cudaStreamSynchronize(strm()));
  return MLEGS_OK;
}

// sum of host doubles over all ranks (MPI_Allreduce of ops:394, 657, 748); no-op on one rank
static int allreduce_host(double *v, int n) {
  Context &c = ctx();
  if (c.nranks == 1) return MLEGS_OK;
  CUDA_TRY(cudaMemcpyAsync(c.d_red, v, n * sizeof(double), cudaMemcpyHostToDevice, strm()));
  MLEGS_TRY(dist_allreduce(c.d_red, n));
  CUDA_TRY(cudaMemcpyAsync(v, c.d_red, n * sizeof(double), cudaMemcpyDeviceToHost, strm()));
  CUDA_TRY(cudaStreamSynchronize(strm()));
  return dist_check_timeout();
}

// log-term coefficients of del^2 P_L_0^0: 4/3, -2, 2/3 over ell^2 exp(lognorm(.,1)) (ops:562-566)
static void ln_del2_coeffs(double c[3]) {
  Context &k = ctx();
  const double ell2 = std::pow(k.p.ell, 2.0);
  c[0] = 4.0 / 3.0 / ell2 / std::exp(k.h_lognorm[0]);
  c[1] = 2.0 / 1.0 / ell2 / std::exp(k.h_lognorm[1]);
  c[2] = 2.0 / 3.0 / ell2 / std::exp(k.h_lognorm[2]);
}

// ---------------------------------------------------------------------------------------------
// masks and filters
// ---------------------------------------------------------------------------------------------
int chop_impl(mlegs_field *s) {
  MLEGS_TRY(ready());
  ChopIdx ci;
  MLEGS_TRY(chop_index(s, &ci));
  MaskArgs a;
  a.e = (cplx *)s->e;
  a.nrl = s->loc_sz[0];
  a.npl = s->loc_sz[1];
  a.nzl = s->loc_sz[2];
  a.r0 = s->loc_st[0];
  a.m0 = s->loc_st[1];
  a.ms = field_mstride(s);
  a.row_mode = s->space[0] == 'F';
  a.nrc = ci.nrc;
  a.npc_rows = ci.npc;
  a.col_cut = (s->space[1] == 'F') ? ci.npc : INT_MAX;
  a.kz_lo = a.kz_hi = 0;
  if (s->space[2] == 'F' && ci.nzc < ci.nzcu) {
    a.kz_lo = ci.nzc;          // 1-based nzc+1 .. nzcu-1
    a.kz_hi = ci.nzcu - 1;
  }
  return launch_mask(a, strm());
}

int dealias_impl(mlegs_field *s) {
  MLEGS_TRY(ready());
  Context &c = ctx();
  MaskArgs a;
  a.e = (cplx *)s->e;
  a.nrl = s->loc_sz[0];
  a.npl = s->loc_sz[1];
  a.nzl = s->loc_sz[2];
  a.r0 = s->loc_st[0];
  a.m0 = s->loc_st[1];
  a.ms = field_mstride(s);
  a.row_mode = 0;
  a.nrc = a.npc_rows = 0;
  a.col_cut = INT_MAX;
  a.kz_lo = a.kz_hi = 0;
  bool any = false;
  if (c.p.np > 1 && s->space[1] == 'F') {
    a.col_cut = std::max(c.p.np / 3 + 1, 1);
    any = true;
  }
  if (c.p.nz > 1 && s->space[2] == 'F') {
    int zcut = std::max(c.p.nz / 3 + 1, 1);
    int zupper = c.p.nz - zcut + 2;
    a.kz_lo = zcut;
    a.kz_hi = zupper - 1;
    any = true;
  }
  if (!any) return MLEGS_OK;
  return launch_mask(a, strm());
}

int svv_impl(mlegs_field *s, double *gain) {
  MLEGS_TRY(ready());
  Context &c = ctx();
  if (!c.p.is_svv) return MLEGS_OK;
  if (!is_space(s, "FFF")) return fail(MLEGS_E_ARG, "svv_filter: scalar must be in FFF space");
  SvvArgs a;
  a.e = (cplx *)s->e;
  a.nrl = s->loc_sz[0];
  a.npl = s->loc_sz[1];
  a.nzl = s->loc_sz[2];
  a.r0 = s->loc_st[0];
  a.m0 = s->loc_st[1];
  a.ms = field_mstride(s);
  a.ak = c.d_ak;
  a.nak = c.p.nz;
  a.qr_den = std::max((double)(c.p.nrchop - 1), 1.0);
  a.qp_den = std::max((double)(c.p.np / 2), 1.0);
  double kmax = 0.0;
  for (double v : c.h_ak) kmax = std::max(kmax, std::fabs(v));
  a.kmax = std::max(kmax, 1.0);
  a.cutoff = std::min(std::max(c.p.svv_cutoff, 0.0), 0.99);
  a.strength = 0.0;
  const double target = std::max(c.p.svv_target, 1.0e-12);
  MLEGS_TRY(launch_svv_energy(a, c.d_red + 16, c.d_red, strm()));
  MLEGS_TRY(dist_allreduce(c.d_red, 2));   // MPI_Allreduce of ops:117-118
  CUDA_TRY(cudaMemcpyAsync(c.h_red, c.d_red, 2 * sizeof(double), cudaMemcpyDeviceToHost, strm()));
  CUDA_TRY(cudaStreamSynchronize(strm()));
  double total = c.h_red[0], tail = c.h_red[1];
  if (total <= 2.2250738585072014e-308) return MLEGS_OK;
  double tail_ratio = tail / total;
  double feedback = std::min(std::max(tail_ratio / target - 1.0, 0.0), 1.0);
  double relax = std::min(std::max(c.p.svv_relax, 0.0), 1.0);
  *gain = (1.0 - relax) * (*gain) + relax * feedback;
  double strength = std::min(std::max(c.p.svv_strength, 0.0) * (*gain), 1.0);
  if (strength <= 0.0) return MLEGS_OK;
  a.strength = strength;
  return launch_svv_apply(a, strm());
}

// calc(k) = sum_n st%e(n, m=0, k) at(n) on the FFF image of s (ops:237-309); result left in d_red
static int calcat_device(const mlegs_field *s, const double *d_at, bool subtract_inplace, mlegs_field *inplace) {
  Context &c = ctx();
  ChopIdx ci;
  MLEGS_TRY(chop_index(s, &ci));
  const mlegs_field *src = s;
  mlegs_field st;
  if (!is_space(s, "FFF")) {
    st = temp_like(s, 5);
    MLEGS_TRY(copy_data(&st, s));
    MLEGS_TRY(trans_impl(&st, "FFF"));
    src = &st;
  }
  if (subtract_inplace) src = inplace;
  int nrows = std::min(src->loc_sz[0], ci.nrc - src->loc_st[0]);
  nrows = std::min(nrows, c.p.nrchop);
  if (2 * src->loc_sz[2] > Context::RED_DOUBLES) return fail(MLEGS_E_ARG, "calcat: nz too large for the workspace");
  if (owns_m0(src)) {
    MLEGS_TRY(launch_calcat((cplx *)src->e, src->loc_sz[0], src->loc_sz[1], src->loc_sz[2], nrows, d_at,
                            (cplx *)c.d_red, subtract_inplace ? 1 : 0, c.h_at1[0], strm()));
  } else {
    CUDA_TRY(cudaMemsetAsync(c.d_red, 0, 2 * src->loc_sz[2] * sizeof(double), strm()));
  }
  return MLEGS_OK;
}

static int calcat_host(const mlegs_field *s, const double *d_at, double *out) {
  Context &c = ctx();
  MLEGS_TRY(calcat_device(s, d_at, false, nullptr));
  MLEGS_TRY(dist_allreduce(c.d_red, 2 * s->loc_sz[2]));   // MPI_Allreduce of ops:265, 302
  CUDA_TRY(cudaMemcpyAsync(out, c.d_red, 2 * s->loc_sz[2] * sizeof(double), cudaMemcpyDeviceToHost, strm()));
  CUDA_TRY(cudaStreamSynchronize(strm()));
  return MLEGS_OK;
}

int zeroat1_impl(mlegs_field *s) {
  MLEGS_TRY(ready());
  MLEGS_TRY(require_fff(s));
  return calcat_device(s, ctx().d_at1, true, s);
}

// ---------------------------------------------------------------------------------------------
// diagonal operators
// ---------------------------------------------------------------------------------------------
int delsqp_impl(mlegs_field *s, bool inverse);

// fftreat, ops:1002-1063: delsqp, radial synthesis of the FFF array ('PFF'), far-field smoothing of every retained
// (m,k) line, radial analysis, idelsqp, zeroat1.
int fftreat_impl(mlegs_field *s) {
  MLEGS_TRY(ready());
  Context &c = ctx();
  ChopIdx ci;
  MLEGS_TRY(chop_index(s, &ci));
  MLEGS_TRY(require_fff(s));
  const double ln = s->ln;
  s->ln = 0.0;                                   // so%ln = 0 (ops:1018); restored below (ops:1058)
  MLEGS_TRY(delsqp_impl(s, false));
  cplx *home = (cplx *)s->e, *tmp = (cplx *)c.d_scratch[0];
  MLEGS_TRY(stage_r(s, false, home, tmp));       // rtrans_backward on the FFF array; s->ln == 0: no log term
  FftreatArgs a;
  a.e = tmp;
  a.nrl = s->loc_sz[0];
  a.npl = s->loc_sz[1];
  a.nzl = s->loc_sz[2];
  a.m0 = s->loc_st[1];
  a.ms = field_mstride(s);
  a.nr = c.p.nr;
  a.ns = c.p.nr * 3 / 4;
  a.ns0 = std::min(a.ns + 4, c.p.nr);
  a.npc = ci.npc;
  a.nzc = ci.nzc;
  a.nzcu = ci.nzcu;
  a.x = c.d_x;
  MLEGS_TRY(launch_fftreat_tail(a, strm()));
  MLEGS_TRY(stage_r(s, true, tmp, home));        // rtrans_forward
  MLEGS_TRY(delsqp_impl(s, true));
  s->ln = ln;
  return zeroat1_impl(s);
}

int delsqp_impl(mlegs_field *s, bool inverse) {
  MLEGS_TRY(ready());
  Context &c = ctx();
  ChopIdx ci;
  MLEGS_TRY(chop_index(s, &ci));
  MLEGS_TRY(require_fff(s));
  const double ell2 = std::pow(c.p.ell, 2.0);
  const bool own = owns_m0(s) && s->loc_st[0] == 0;
  double ln_new = 0.0;
  if (inverse && own) {
    cplx v;
    MLEGS_TRY(read_elem(s, 0, &v));
    ln_new = v.x * ell2 * std::exp(c.h_lognorm[0]);   // ops:392
  }
  if (inverse) MLEGS_TRY(allreduce_host(&ln_new, 1));   // ops:394
  MLEGS_TRY(launch_delsqp((cplx *)s->e, s->loc_sz[0], s->loc_sz[1], s->loc_sz[2], s->loc_st[1], field_mstride(s), ci.nrc, ci.npc,
                          ell2, inverse ? 1 : 0, strm()));
  if (own) {
    if (!inverse) {
      PokeArgs p{};
      p.n = 1;
      p.off[0] = 0;
      p.re[0] = s->ln / ell2 / std::exp(c.h_lognorm[0]);   // ops:359
      p.im[0] = 0.0;
      p.mode[0] = 0;
      MLEGS_TRY(launch_poke((cplx *)s->e, p, strm()));
    } else {
      MLEGS_TRY(launch_zero_line((cplx *)s->e, 0, (long long)s->loc_sz[0] * s->loc_sz[1], s->loc_sz[2], strm()));
    }
  }
  s->ln = inverse ? ln_new : 0.0;
  return MLEGS_OK;
}

// ---------------------------------------------------------------------------------------------
// banded operators
// ---------------------------------------------------------------------------------------------
static int band_args(const mlegs_field *s, BandOpArgs *a) {
  Context &c = ctx();
  ChopIdx ci;
  MLEGS_TRY(chop_index(s, &ci));
  a->e = (cplx *)s->e;
  a->nrl = s->loc_sz[0];
  a->npl = s->loc_sz[1];
  a->nzl = s->loc_sz[2];
  a->m0 = s->loc_st[1];
  a->ms = field_mstride(s);
  a->ne = c.ne;
  a->nrc = ci.nrc;
  a->npc = ci.npc;
  a->nzc = ci.nzc;
  a->nzcu = ci.nzcu;
  a->napply = 1;
  a->combine = 0;
  a->alpha = a->beta = 0.0;
  a->nlnc = 0;
  a->lnc[0] = a->lnc[1] = a->lnc[2] = 0.0;
  if (ci.nrc + 2 > c.ne) return fail(MLEGS_E_ARG, "band operator: chopping in r exceeds the normalisation table");
  return MLEGS_OK;
}

// src != nullptr (xxdx, del2): the operand is read from there (same layout and metadata as s) and the result written to
// s%e -- `t = s; call op(t)` without the copy.  negate: the whole result is negated on the way out.
int xxdx_impl(mlegs_field *s, const void *src = nullptr);
int del2_impl(mlegs_field *s, bool horizontal, const void *src = nullptr, bool negate = false);

int xxdx_impl(mlegs_field *s, const void *src) {
  MLEGS_TRY(ready());
  Context &c = ctx();
  if (!src) MLEGS_TRY(require_fff(s));
  BandOpArgs a;
  MLEGS_TRY(band_args(s, &a));
  a.src = (const cplx *)src;
  a.tab = c.d_xxdx;
  a.nb = 3;
  a.ak = nullptr;
  if (s->ln != 0.0) {   // ops:453-456
    a.nlnc = 2;
    a.lnc[0] = 1.0 / std::exp(c.h_lognorm[0]) * s->ln;
    a.lnc[1] = 1.0 / std::exp(c.h_lognorm[1]) * s->ln;
  }
  s->ln = 0.0;
  return launch_band_op(a, strm());
}

static void set_del2_ln(BandOpArgs *a, double ln) {
  if (ln == 0.0) return;
  double cf[3];
  ln_del2_coeffs(cf);
  a->nlnc = 3;
  a->lnc[0] = cf[0] * ln;
  a->lnc[1] = -(cf[1] * ln);
  a->lnc[2] = cf[2] * ln;
}

int del2_impl(mlegs_field *s, bool horizontal, const void *src, bool negate) {
  MLEGS_TRY(ready());
  Context &c = ctx();
  if (!src) MLEGS_TRY(require_fff(s));
  BandOpArgs a;
  MLEGS_TRY(band_args(s, &a));
  a.src = (const cplx *)src;
  a.out_neg = negate ? 1 : 0;
  a.tab = c.d_del2h;
  a.nb = 5;
  a.ak = horizontal ? nullptr : c.d_ak;
  set_del2_ln(&a, s->ln);
  s->ln = 0.0;
  return launch_band_op(a, strm());
}

int helmp_impl(mlegs_field *s, int power, double alpha, double beta, const void *src = nullptr);
// src != nullptr: the operand is read from there (same layout as s) and the result written to s%e
int helmp_impl(mlegs_field *s, int power, double alpha, double beta, const void *src) {
  MLEGS_TRY(ready());
  Context &c = ctx();
  if (!((power % 2 == 0) && power >= 4)) return fail(MLEGS_E_ARG, "helmp: even power greater than or equal to 4");
  if (power > 8)
    return fail(MLEGS_E_ARG, "helmp: power must be less than or equal to 8 (supported power = 4, 6 or 8)");
  MLEGS_TRY(require_fff(s));
  BandOpArgs a;
  MLEGS_TRY(band_args(s, &a));
  a.tab = c.d_del2h;
  a.nb = 5;
  a.ak = c.d_ak;
  a.napply = power / 2;
  a.combine = 1;
  a.src = (const cplx *)src;
  a.alpha = alpha;
  a.beta = beta;
  set_del2_ln(&a, s->ln);
  s->ln = alpha * s->ln;
  return launch_band_op(a, strm());
}

// the two axial loops of the solves (e.g. ops:828-841): planes [0, min(nzl,nzc)) then [max(nzcu,1)-1, nzl)
static int solve_two_ranges(SolveArgs base, const ChopIdx &ci, int nzl, int kl_first) {
  const int n1 = std::min(nzl, ci.nzc);
  const int lo = std::max(ci.nzcu, 1) - 1;
  return launch_band_solve_ranges(base, n1, kl_first, lo, nzl, strm());
}

static int solve_args(const mlegs_field *s, const ChopIdx &ci, SolveArgs *a) {
  Context &c = ctx();
  memset(a, 0, sizeof(*a));
  a->e = (cplx *)s->e;
  a->nrl = s->loc_sz[0];
  a->npl = s->loc_sz[1];
  a->m0 = s->loc_st[1];
  a->ms = field_mstride(s);
  a->tab = c.d_del2h;
  a->ne = c.ne;
  a->ak = c.d_ak;
  a->nrc = ci.nrc;
  a->npc = ci.npc;
  a->nnmax = std::max(ci.nrc, 1);
  a->kl = a->ku = 2;
  a->power = 2;
  if (ci.nrc + 2 > c.ne) return fail(MLEGS_E_ARG, "band solve: chopping in r exceeds the normalisation table");
  return MLEGS_OK;
}

int ihelm_impl(mlegs_field *s, double alpha) {
  MLEGS_TRY(ready());
  ChopIdx ci;
  MLEGS_TRY(chop_index(s, &ci));
  MLEGS_TRY(require_fff(s));
  if (std::fabs(alpha) < 5.0e-14)
    return fail(MLEGS_E_ARG, "ihelm: alpha equals to zero. Inversion of 0*identity is impossible");
  s->ln = s->ln / alpha;
  if (s->ln != 0.0 && owns_m0(s)) {   // ops:820-824
    double cf[3];
    ln_del2_coeffs(cf);
    PokeArgs p{};
    p.n = 3;
    for (int i = 0; i < 3; ++i) {
      p.off[i] = i;
      p.im[i] = 0.0;
      p.mode[i] = 1;
    }
    p.re[0] = -(cf[0] * s->ln);
    p.re[1] = cf[1] * s->ln;
    p.re[2] = -(cf[2] * s->ln);
    MLEGS_TRY(launch_poke((cplx *)s->e, p, strm()));
  }
  SolveArgs a;
  MLEGS_TRY(solve_args(s, ci, &a));
  a.add_alpha = 1;
  a.alpha = alpha;
  return solve_two_ranges(a, ci, s->loc_sz[2], 2);
}

int idel2_impl(mlegs_field *s, int have_preln, double preln) {
  MLEGS_TRY(ready());
  Context &c = ctx();
  ChopIdx ci;
  MLEGS_TRY(chop_index(s, &ci));
  MLEGS_TRY(require_fff(s));
  const double ell2 = std::pow(c.p.ell, 2.0);
  SolveArgs a;
  MLEGS_TRY(solve_args(s, ci, &a));
  a.special00 = have_preln ? 2 : 1;
  a.sp0 = 4.0 / 3.0 / ell2;
  a.sp1 = 2.0 / 1.0 / ell2 * std::exp(c.h_lognorm[0] - c.h_lognorm[1]);
  a.sp2 = 2.0 / 3.0 / ell2 * std::exp(c.h_lognorm[0] - c.h_lognorm[2]);
  a.preln_rhs = preln / std::exp(c.h_lognorm[0]);
  MLEGS_TRY(solve_two_ranges(a, ci, s->loc_sz[2], have_preln ? 3 : 2));
  if (ci.npc >= 1 && ci.nrc >= 1 && ci.nzc >= 1) {
    double lnv = 0.0;
    if (owns_m0(s)) {
      cplx v;
      MLEGS_TRY(read_elem(s, 0, &v));
      lnv = v.x * std::exp(c.h_lognorm[0]);   // ops:625, 716
    }
    MLEGS_TRY(allreduce_host(&lnv, 1));       // ops:657, 748
    s->ln = lnv;
  }
  return MLEGS_OK;
}

int ihelmp_impl(mlegs_field *s, int power, double alpha, double beta) {
  MLEGS_TRY(ready());
  Context &c = ctx();
  if (!((power % 2 == 0) && power >= 4)) return fail(MLEGS_E_ARG, "ihelmp: even power greater than or equal to 4");
  if (power > 8)
    return fail(MLEGS_E_ARG, "ihelmp: power must be less than or equal to 8 (supported power = 4, 6 or 8)");
  ChopIdx ci;
  MLEGS_TRY(chop_index(s, &ci));
  MLEGS_TRY(require_fff(s));
  if (std::fabs(alpha) < 5.0e-14)
    return fail(MLEGS_E_ARG, "ihelmp: alpha equals to zero. Inversion of 0*identity is impossible");
  s->ln = s->ln / alpha;
  if (s->ln != 0.0 && owns_m0(s)) {
    // ops:939-951: bl = del2^(p/2-1) bl2 on the (m=0,k=0) column, subtracted from the right-hand side
    const int nrc = ci.nrc;
    double cf[3];
    ln_del2_coeffs(cf);
    std::vector<double> bl2(nrc, 0.0), bl(nrc), tmp(nrc);
    if (nrc > 0) bl2[0] = cf[0] * s->ln;
    if (nrc > 1) bl2[1] = -cf[1] * s->ln;
    if (nrc > 2) bl2[2] = cf[2] * s->ln;
    bl = bl2;
    const double *t5 = c.h_del2h.data();   // m = 0
    const double ak2 = c.h_ak[0] * c.h_ak[0];
    for (int q = 0; q < power / 2 - 1; ++q) {
      for (int i = 0; i < nrc; ++i) {
        double acc = 0.0;
        for (int b = 0; b < 5; ++b) {
          int j = i + b - 2;
          if (j < 0 || j >= nrc) continue;
          double cfv = t5[(size_t)b * c.ne + i];
          if (b == 2) cfv = cfv - ak2;
          acc = acc + bl[j] * cfv;
        }
        tmp[i] = acc;
      }
      bl = tmp;
    }
    std::vector<double> bb(nrc);
    for (int i = 0; i < nrc; ++i) bb[i] = beta * bl2[i];
    if (2 * nrc > Context::RED_DOUBLES) return fail(MLEGS_E_ARG, "ihelmp: nrc too large for the workspace");
    memcpy(c.h_red, bl.data(), nrc * sizeof(double));
    memcpy(c.h_red + nrc, bb.data(), nrc * sizeof(double));
    CUDA_TRY(cudaMemcpyAsync(c.d_red, c.h_red, 2 * nrc * sizeof(double), cudaMemcpyHostToDevice, strm()));
    MLEGS_TRY(launch_col_update((cplx *)s->e, std::min(nrc, s->loc_sz[0]), 0, c.d_red, c.d_red + nrc, 0.0, strm()));
    CUDA_TRY(cudaStreamSynchronize(strm()));   // h_red is reused by later calls
  }
  SolveArgs a;
  MLEGS_TRY(solve_args(s, ci, &a));
  a.kl = a.ku = power;
  a.power = power;
  a.add_alpha = 1;
  a.alpha = alpha;
  a.beta = beta;
  return solve_two_ranges(a, ci, s->loc_sz[2], power);
}

// ---------------------------------------------------------------------------------------------
// time integrators
// ---------------------------------------------------------------------------------------------
static double hv_signed() {
  const mlegs_params &p = ctx().p;
  return p.hypervisc * std::pow(-1.0, p.hyperpow / 2 + 1);
}

static int check_fff(const mlegs_field *a) {
  if (!is_space(a, "FFF")) return fail(MLEGS_E_ARG, "fefe: all input scalars must be in FFF for time stepping");
  return MLEGS_OK;
}

static int lin(int mode, mlegs_field *y, const mlegs_field *x1, const mlegs_field *x2, const mlegs_field *x3, double a,
               double b, double cc) {
  LinArgs p;
  p.mode = mode;
  p.n = nelem(y);
  p.y = (cplx *)y->e;
  p.x1 = x1 ? (const cplx *)x1->e : nullptr;
  p.x2 = x2 ? (const cplx *)x2->e : nullptr;
  p.x3 = x3 ? (const cplx *)x3->e : nullptr;
  p.a = a;
  p.b = b;
  p.c = cc;
  return launch_lincomb(p, strm());
}

// svis of fefe/abab/abcn (e.g. ops:1217-1230): returns the un-scaled operator image in `svis` and the factor
static int viscous_term(const mlegs_field *s, mlegs_field *svis, double *factor, bool *zero) {
  const mlegs_params &p = ctx().p;
  *zero = false;
  if (p.hyperpow != 0 && is_space(s, "FFF")) {
    // svis = helmp(s) without the copy svis = s: the band operator reads s%e and writes svis%e
    void *e = svis->e;
    *svis = *s;
    svis->e = e;
    MLEGS_TRY(helmp_impl(svis, p.hyperpow, 0.0, p.visc / hv_signed(), s->e));
    *factor = hv_signed();
    return MLEGS_OK;
  }
  MLEGS_TRY(copy_data(svis, s));
  if (p.hyperpow == 0) {
    if (p.visc < 5.0e-14) {
      *zero = true;
      *factor = 0.0;
      CUDA_TRY(cudaMemsetAsync(svis->e, 0, nelem(svis) * sizeof(cplx), strm()));
    } else {
      MLEGS_TRY(del2_impl(svis, false));
      *factor = p.visc;
    }
  } else {
    MLEGS_TRY(helmp_impl(svis, p.hyperpow, 0.0, p.visc / hv_signed()));
    *factor = hv_signed();
  }
  return MLEGS_OK;
}

int fefe_impl(mlegs_field *s, const mlegs_field *nl, double dt) {
  MLEGS_TRY(ready());
  MLEGS_TRY(check_fff(s));
  MLEGS_TRY(check_fff(nl));
  mlegs_field svis = temp_like(s, 2);
  double fac;
  bool zero;
  MLEGS_TRY(viscous_term(s, &svis, &fac, &zero));
  MLEGS_TRY(lin(7, s, nl, &svis, nullptr, dt, 0.0, zero ? 1.0 : fac));
  s->ln = s->ln + dt * (nl->ln + svis.ln);
  return MLEGS_OK;
}

int febe_impl(mlegs_field *s, const mlegs_field *nl, double dt) {
  MLEGS_TRY(ready());
  MLEGS_TRY(check_fff(s));
  MLEGS_TRY(check_fff(nl));
  const mlegs_params &p = ctx().p;
  mlegs_field sh = temp_like(s, 1);
  MLEGS_TRY(lin(2, &sh, s, nl, nullptr, dt, 0.0, 0.0));
  sh.ln = s->ln + dt * nl->ln;
  if (p.hyperpow == 0) {
    if (p.visc < 5.0e-14)
      return fail(MLEGS_E_ARG,
                  "febe: inviscid case and no linear term in rhs. semi-implicit time adv is impossible");
    double a = -1.0 / (dt * p.visc);
    sh.ln = a * sh.ln;
    MLEGS_TRY(lin(4, &sh, nullptr, nullptr, nullptr, a, 0.0, 0.0));
    MLEGS_TRY(ihelm_impl(&sh, a));
  } else {
    double a = -1.0 / (dt * hv_signed());
    double b = p.visc / hv_signed();
    sh.ln = a * sh.ln;
    MLEGS_TRY(lin(4, &sh, nullptr, nullptr, nullptr, a, 0.0, 0.0));
    MLEGS_TRY(ihelmp_impl(&sh, p.hyperpow, a, b));
  }
  return copy_data(s, &sh);
}

int abcn_impl(mlegs_field *s, mlegs_field *s_p, mlegs_field *nl, mlegs_field *nl_p, double dt) {
  MLEGS_TRY(ready());
  MLEGS_TRY(check_fff(s));
  MLEGS_TRY(check_fff(nl));
  MLEGS_TRY(check_fff(s_p));
  MLEGS_TRY(check_fff(nl_p));
  const mlegs_params &p = ctx().p;
  if (p.hyperpow == 0 && p.visc < 5.0e-14)
    return fail(MLEGS_E_ARG, "abcn: inviscid case and no linear term in rhs. semi-implicit time adv is impossible");
  // The reference builds sh = s + dt (1.5 nl - 0.5 nl_p), svis = L s, sh = a (sh + dt/2 svis), solves, and assigns
  // s = sh (ops:1214-1255).  Here svis is computed first (from the old s), then ONE pass writes a (sh + dt/2 svis)
  // straight into s%e (same operations in the same order, the intermediate sh rounded to double as the reference
  // stores it), and the solve runs in place: two field copies and one full pass fewer per call.
  const double ln_sh = s->ln + dt * (1.5 * nl->ln - 0.5 * nl_p->ln);
  mlegs_field svis = temp_like(s, 2);
  double fac;
  bool zero;
  MLEGS_TRY(viscous_term(s, &svis, &fac, &zero));
  const double a = (p.hyperpow == 0) ? -2.0 / (dt * p.visc) : -2.0 / (dt * hv_signed());
  {
    LinArgs q;
    q.mode = 9;
    q.n = nelem(s);
    q.y = (cplx *)s->e;
    q.x1 = (const cplx *)nl->e;
    q.x2 = (const cplx *)nl_p->e;
    q.x3 = (const cplx *)svis.e;
    q.a = a;
    q.b = dt / 2.0;
    q.c = fac;
    q.d = dt;
    MLEGS_TRY(launch_lincomb(q, strm()));
  }
  s->ln = a * (ln_sh + dt / 2.0 * svis.ln);
  if (p.hyperpow == 0)
    MLEGS_TRY(ihelm_impl(s, a));
  else
    MLEGS_TRY(ihelmp_impl(s, p.hyperpow, a, p.visc / hv_signed()));
  MLEGS_TRY(copy_data(s_p, s));      // ops:1256: s_p receives the NEW s
  MLEGS_TRY(copy_data(nl_p, nl));
  return MLEGS_OK;
}

// abab, ops:1096-1155.  Quirk Q2 reproduced: the inviscid branch zeroes `svis` twice (ops:1135) and leaves svis_p a
// plain copy of s_p.
int abab_impl(mlegs_field *s, mlegs_field *s_p, mlegs_field *nl, mlegs_field *nl_p, double dt, int is_2nd_svis_p) {
  MLEGS_TRY(ready());
  MLEGS_TRY(check_fff(s));
  MLEGS_TRY(check_fff(nl));
  MLEGS_TRY(check_fff(s_p));
  MLEGS_TRY(check_fff(nl_p));
  const mlegs_params &p = ctx().p;
  mlegs_field svis = temp_like(s, 2), svis_p = temp_like(s_p, 3);
  double fac, fac_p = 1.0;
  bool zero, zero_p = false;
  MLEGS_TRY(viscous_term(s, &svis, &fac, &zero));
  const bool inviscid = (p.hyperpow == 0 && p.visc < 5.0e-14);
  if (is_2nd_svis_p || inviscid) {
    MLEGS_TRY(copy_data(&svis_p, s_p));
  } else {
    MLEGS_TRY(viscous_term(s_p, &svis_p, &fac_p, &zero_p));
  }
  LinArgs a;
  a.mode = 8;
  a.n = nelem(s);
  a.y = (cplx *)s->e;
  a.x1 = (const cplx *)nl->e;
  a.x2 = (const cplx *)svis.e;
  a.x3 = (const cplx *)nl_p->e;
  a.x4 = (const cplx *)svis_p.e;
  a.a = dt;
  a.b = 0.0;
  a.c = zero ? 1.0 : fac;
  a.d = fac_p;
  MLEGS_TRY(launch_lincomb(a, strm()));
  s->ln = s->ln + dt * (1.5 * (nl->ln + svis.ln) - 0.5 * (nl_p->ln + svis_p.ln));
  MLEGS_TRY(copy_data(s_p, s));
  MLEGS_TRY(copy_data(nl_p, nl));
  return MLEGS_OK;
}

// helm, ops:762-789: s <- del2(s) + alpha*s, ln <- alpha*ln.  The reference forms this in a local scalar and never
// returns it (its helm leaves s unchanged); the documented operator is what is provided here.
int helm_impl(mlegs_field *s, double alpha) {
  MLEGS_TRY(ready());
  MLEGS_TRY(require_fff(s));
  mlegs_field so = temp_like(s, 2);
  MLEGS_TRY(copy_data(&so, s));
  MLEGS_TRY(del2_impl(&so, false));
  const double ln = alpha * s->ln;
  MLEGS_TRY(lin(0, s, &so, nullptr, nullptr, 1.0, alpha, 0.0));   // s = so + alpha*s
  s->ln = ln;
  return MLEGS_OK;
}

// ---------------------------------------------------------------------------------------------
// vector-field operations
// ---------------------------------------------------------------------------------------------
int vecprod_impl(mlegs_field *vr, mlegs_field *vp, mlegs_field *vz, const mlegs_field *ur, const mlegs_field *up,
                 const mlegs_field *uz) {
  MLEGS_TRY(ready());
  Context &c = ctx();
  const char *names[6] = {"vr", "vp", "vz", "ur", "up", "uz"};
  const mlegs_field *f[6] = {vr, vp, vz, ur, up, uz};
  for (int i = 0; i < 6; ++i)
    if (!is_space(f[i], "PPP"))
      return fail(MLEGS_E_ARG, std::string("vector_product: to compute v x u, ") + names[i] + " must be in PPP");
  return launch_vecprod((cplx *)vr->e, (cplx *)vp->e, (cplx *)vz->e, (const cplx *)ur->e, (const cplx *)up->e,
                        (const cplx *)uz->e, vr->loc_sz[0], vr->loc_sz[1], vr->loc_sz[2], vr->loc_st[0], c.p.nr,
                        c.p.np / 2, c.p.nz, strm());
}

int trans_many_scaled(int n, mlegs_field *const *s, const char *to, unsigned rs_mask, const void *const *src);   // ops_trans.cu

// rows i < nr of a PPP scalar times r(i) (divide = 0) or over r(i) (divide = 1)
int launch_rscale_field(mlegs_field *f, int divide) {
  Context &c = ctx();
  const size_t ncols = (size_t)f->loc_sz[1] * f->loc_sz[2];
  return launch_rscale((cplx *)f->e, f->loc_sz[0], ncols, f->loc_st[0], c.p.nr, c.d_r, divide, strm());
}

int tp2vec_impl(const mlegs_field *psi, const mlegs_field *chi, mlegs_field *vr, mlegs_field *vp, mlegs_field *vz) {
  MLEGS_TRY(ready());
  Context &c = ctx();
  if (!is_space(psi, "FFF")) return fail(MLEGS_E_ARG, "vector_projection: psi must be in FFF");
  if (!is_space(chi, "FFF")) return fail(MLEGS_E_ARG, "vector_projection: chi must be in FFF");
  if (!is_space(vr, "PPP")) return fail(MLEGS_E_ARG, "vector_projection: vr must be in PPP");
  if (!is_space(vp, "PPP")) return fail(MLEGS_E_ARG, "vector_projection: vp must be in PPP");
  if (!is_space(vz, "PPP")) return fail(MLEGS_E_ARG, "vector_projection: vz must be in PPP");
  // ur = chi, up = psi, uz = chi with nrchop offset 3 (ops:1479-1481); they live in the output buffers.  The copies
  // are not made: the operators below read chi / psi and write into the output buffers (metadata copied here).
  auto take_meta = [](mlegs_field *dst, const mlegs_field *src) {
    void *e = dst->e;
    *dst = *src;
    dst->e = e;
  };
  take_meta(vr, chi);
  take_meta(vp, psi);
  take_meta(vz, chi);
  mlegs_field *ur = vr, *up = vp, *uz = vz;
  ur->nrchop_offset = up->nrchop_offset = uz->nrchop_offset = 3;
  ur->npchop_offset = up->npchop_offset = uz->npchop_offset = 0;
  ur->nzchop_offset = up->nzchop_offset = uz->nzchop_offset = 0;
  ChopIdx ci;
  MLEGS_TRY(chop_index(ur, &ci));
  MLEGS_TRY(xxdx_impl(ur, chi->e));
  MLEGS_TRY(xxdx_impl(up, psi->e));
  TvCombineArgs t;
  t.ur = (cplx *)ur->e;
  t.up = (cplx *)up->e;
  t.psi = (const cplx *)psi->e;
  t.uz = (const cplx *)chi->e;               // uz still equals chi at this point (ops:1481, 1494)
  t.nrl = ur->loc_sz[0];
  t.npl = ur->loc_sz[1];
  t.nzl = ur->loc_sz[2];
  t.m0 = ur->loc_st[1];
  t.ms = field_mstride(ur);
  t.nrc = ci.nrc;
  t.npc = ci.npc;
  t.nzc = ci.nzc;
  t.nzcu = ci.nzcu;
  t.ak = c.d_ak;
  MLEGS_TRY(launch_tv_combine(t, strm()));
  ur->nrchop_offset = up->nrchop_offset = 0;
  MLEGS_TRY(chop_impl(ur));
  MLEGS_TRY(chop_impl(up));
  MLEGS_TRY(del2_impl(uz, true, chi->e, true));      // uz = -del2h(chi) (ops:1488-1489), one pass
  uz->nrchop_offset = 0;
  MLEGS_TRY(chop_impl(uz));
  // the three backward transforms (ops:1503-1505, 1541) are independent: one launch per stage for all of them
  // ur/r and up/r (ops:1509-1527) ride on the stores of the backward azimuthal FFT
  mlegs_field *comps[3] = {ur, up, uz};
  return trans_many_scaled(3, comps, "PPP", 0x3u, nullptr);
}

int tp2curlvec_impl(const mlegs_field *psi, const mlegs_field *chi, mlegs_field *wr, mlegs_field *wp, mlegs_field *wz) {
  MLEGS_TRY(ready());
  mlegs_field mdel2chi = temp_like(chi, 3);           // metadata of chi, data in a scratch buffer
  if (is_space(chi, "FFF")) {
    MLEGS_TRY(del2_impl(&mdel2chi, false, chi->e, true));   // -del2(chi) (ops:1553-1554) without the copy, one pass
  } else {                                            // del2 transforms its operand first (ops:531-535)
    MLEGS_TRY(copy_data(&mdel2chi, chi));
    MLEGS_TRY(del2_impl(&mdel2chi, false, nullptr, true));
  }
  return tp2vec_impl(&mdel2chi, psi, wr, wp, wz);
}

int vec2tp_impl(const mlegs_field *vr, const mlegs_field *vp, const mlegs_field *vz, mlegs_field *psi,
                mlegs_field *chi) {
  MLEGS_TRY(ready());
  Context &c = ctx();
  if (!is_space(vr, "PPP")) return fail(MLEGS_E_ARG, "vector_projection: vr must be in PPP");
  if (!is_space(vp, "PPP")) return fail(MLEGS_E_ARG, "vector_projection: vr must be in PPP");
  if (!is_space(vz, "PPP")) return fail(MLEGS_E_ARG, "vector_projection: vz must be in PPP");
  if (!is_space(psi, "FFF")) return fail(MLEGS_E_ARG, "vector_projection: psi must be in FFF");
  if (!is_space(chi, "FFF")) return fail(MLEGS_E_ARG, "vector_projection: chi must be in FFF");
  const bool has_z = c.p.nz > 1;
  CUDA_TRY(cudaMemsetAsync(psi->e, 0, nelem(psi) * sizeof(cplx), strm()));
  CUDA_TRY(cudaMemsetAsync(chi->e, 0, nelem(chi) * sizeof(cplx), strm()));
  psi->ln = 0.0;
  chi->ln = 0.0;
  mlegs_field ur = temp_like(vr, 1), up = temp_like(vp, 2), uz = temp_like(vz, 3);
  {   // metadata of the inputs; the data is read straight from vr, vp, vz by the first transform stage
    void *e1 = ur.e, *e2 = up.e, *e3 = uz.e;
    ur = *vr;
    up = *vp;
    uz = *vz;
    ur.e = e1;
    up.e = e2;
    uz.e = e3;
  }

  // ur, uz -> 'PFF' (phi and z spectral, r physical), ops:1357-1363, 1375-1381
  {   // the three azimuthal FFTs in one launch (r*ur, r*up of ops:1337-1355 fused into its loads, no copies of the
      // inputs), then the axial FFTs of ur and uz in one launch
    mlegs_field *comps[3] = {&ur, &up, &uz};
    const void *srcs[3] = {vr->e, vp->e, vz->e};
    MLEGS_TRY(trans_many_scaled(3, comps, "PFP", 0x3u, srcs));
    if (has_z) {
      FieldBatch fb;
      fb.n = 2;
      fb.in[0] = fb.out[0] = (cplx *)ur.e;
      fb.in[1] = fb.out[1] = (cplx *)uz.e;
      MLEGS_TRY(stage_z(&ur, true, nullptr, nullptr, &fb));
    }
  }
  set_space3(&ur, "PFF");
  set_space3(&uz, "PFF");
  // up -> FFF, far-field value, back to 'PFF' (ops:1365-1373)
  MLEGS_TRY(trans_impl(&up, "FFF"));
  std::vector<double> inf(2 * (size_t)up.loc_sz[2]);
  MLEGS_TRY(calcat_host(&up, c.d_at1, inf.data()));
  psi->ln = -1.0 / 2.0 * inf[0];
  {
    cplx *dst = (cplx *)c.d_scratch[4];
    MLEGS_TRY(stage_r(&up, false, (cplx *)up.e, dst));
    up.e = dst;   // up now lives in scratch 4; scratch 2 is free
  }
  set_space3(&up, "PFF");
  if (owns_m0(&up))
    MLEGS_TRY(launch_col_update((cplx *)up.e, c.p.nr, 1, c.d_x, nullptr, psi->ln, strm()));

  ur.nrchop_offset = up.nrchop_offset = uz.nrchop_offset = 3;
  ChopIdx ci;
  MLEGS_TRY(chop_index(&ur, &ci));
  if (ci.nrc + 1 > c.ne) return fail(MLEGS_E_ARG, "vector_projection: chopping in r exceeds the table");

  // five parity-folded contractions per (m,k) (ops:1413-1435) as five batched GEMMs
  LegArgs g;
  g.w = nullptr;
  g.lnx = c.d_lnx;
  g.nr = c.p.nr;
  g.nrh = c.nrh;
  g.ne = c.ne;
  g.nrl = ur.loc_sz[0];
  g.npl = ur.loc_sz[1];
  g.m0 = ur.loc_st[1];
  g.ms = field_mstride(&ur);
  g.nzl = ur.loc_sz[2];
  g.nrc = ci.nrc;
  g.npc = ci.npc;
  g.nrdim = c.nrdim;
  g.lnval = 0.0;
  g.peer = nullptr;
  g.npdim = c.npdim;
  cplx *T = (cplx *)c.d_scratch[2];
  g.out = T;
  TpCombineArgs t;
  t.t = T;
  t.nrl = g.nrl;
  t.npl = g.npl;
  t.nzl = g.nzl;
  t.m0 = g.m0;
  t.ms = g.ms;
  t.nrc = ci.nrc;
  t.npc = ci.npc;
  t.nzc = ci.nzc;
  t.nzcu = ci.nzcu;
  t.ak = c.d_ak;
  struct Step {
    const double *tab;
    const double *w;
    const cplx *in;
    int swap, skip0, mode;
    cplx *dst;
  } steps[5] = {
      {c.d_vtab, nullptr, (const cplx *)ur.e, 0, 1, 0, (cplx *)psi->e},   // psi  = -iu*mv*eomul(v, ur)
      {c.d_dtab, nullptr, (const cplx *)up.e, 1, 0, 1, (cplx *)psi->e},   // psi -= oemul(d, up)
      {c.d_dtab, nullptr, (const cplx *)ur.e, 1, 0, 2, (cplx *)chi->e},   // chi  = iu*kv*oemul(d, ur)
      {c.d_vtab, nullptr, (const cplx *)up.e, 0, 1, 3, (cplx *)chi->e},   // chi += mv*kv*eomul(v, up)
      {c.d_pf, c.d_w, (const cplx *)uz.e, 0, 0, 4, (cplx *)chi->e},       // chi -= eomul(t, uz), t = pf*w
  };
  for (int q = 0; q < 5; ++q) {
    g.pf = steps[q].tab;
    g.w = steps[q].w;
    g.in = steps[q].in;
    g.swap_parity = steps[q].swap;
    g.skip_m0 = steps[q].skip0;
    MLEGS_TRY(launch_leg_forward(g, strm()));
    t.dst = steps[q].dst;
    t.mode = steps[q].mode;
    MLEGS_TRY(launch_tp_combine(t, strm()));
  }
  MLEGS_TRY(idel2_impl(chi, 0, 0.0));
  psi->nrchop_offset = psi->npchop_offset = psi->nzchop_offset = 0;
  chi->nrchop_offset = chi->npchop_offset = chi->nzchop_offset = 0;
  MLEGS_TRY(chop_impl(psi));
  MLEGS_TRY(chop_impl(chi));
  MLEGS_TRY(zeroat1_impl(psi));
  MLEGS_TRY(zeroat1_impl(chi));
  return MLEGS_OK;
}

}  // namespace mlegs

using namespace mlegs;

extern "C" {

int mlegs_b200_chop(mlegs_field *s) { return chop_impl(s); }
int mlegs_b200_dealias(mlegs_field *s) { return dealias_impl(s); }
int mlegs_b200_svv_filter(mlegs_field *s, double *gain) { return svv_impl(s, gain); }
int mlegs_b200_calcat0(const mlegs_field *s, double *out) {
  MLEGS_TRY(ready());
  return calcat_host(s, ctx().d_at0, out);
}
int mlegs_b200_calcat1(const mlegs_field *s, double *out) {
  MLEGS_TRY(ready());
  return calcat_host(s, ctx().d_at1, out);
}
int mlegs_b200_zeroat1(mlegs_field *s) { return zeroat1_impl(s); }
int mlegs_b200_fftreat(mlegs_field *s) { return fftreat_impl(s); }
int mlegs_b200_delsqp(mlegs_field *s) { return delsqp_impl(s, false); }
int mlegs_b200_idelsqp(mlegs_field *s) { return delsqp_impl(s, true); }
int mlegs_b200_xxdx(mlegs_field *s) { return xxdx_impl(s); }
int mlegs_b200_del2h(mlegs_field *s) { return del2_impl(s, true); }
int mlegs_b200_del2(mlegs_field *s) { return del2_impl(s, false); }
int mlegs_b200_idel2(mlegs_field *s, int have_preln, double preln) { return idel2_impl(s, have_preln, preln); }
int mlegs_b200_ihelm(mlegs_field *s, double alpha) { return ihelm_impl(s, alpha); }
int mlegs_b200_helmp(mlegs_field *s, int power, double alpha, double beta) { return helmp_impl(s, power, alpha, beta); }
int mlegs_b200_ihelmp(mlegs_field *s, int power, double alpha, double beta) {
  return ihelmp_impl(s, power, alpha, beta);
}
int mlegs_b200_solve_cache(int on) {
  CUDA_TRY(cudaDeviceSynchronize());
  band_solve_cache_enable(on);
  return MLEGS_OK;
}
int mlegs_b200_fefe(mlegs_field *s, const mlegs_field *nl, double dt) { return fefe_impl(s, nl, dt); }
int mlegs_b200_febe(mlegs_field *s, const mlegs_field *nl, double dt) { return febe_impl(s, nl, dt); }
int mlegs_b200_abcn(mlegs_field *s, mlegs_field *s_p, mlegs_field *nl, mlegs_field *nl_p, double dt) {
  return abcn_impl(s, s_p, nl, nl_p, dt);
}
int mlegs_b200_abab(mlegs_field *s, mlegs_field *s_p, mlegs_field *nl, mlegs_field *nl_p, double dt,
                    int is_2nd_svis_p) {
  return abab_impl(s, s_p, nl, nl_p, dt, is_2nd_svis_p);
}
int mlegs_b200_helm(mlegs_field *s, double alpha) { return helm_impl(s, alpha); }
int mlegs_b200_vecprod(mlegs_field *vr, mlegs_field *vp, mlegs_field *vz, const mlegs_field *ur,
                       const mlegs_field *up, const mlegs_field *uz) {
  return vecprod_impl(vr, vp, vz, ur, up, uz);
}
int mlegs_b200_vec2tp(const mlegs_field *vr, const mlegs_field *vp, const mlegs_field *vz, mlegs_field *psi,
                      mlegs_field *chi) {
  return vec2tp_impl(vr, vp, vz, psi, chi);
}
int mlegs_b200_tp2vec(const mlegs_field *psi, const mlegs_field *chi, mlegs_field *vr, mlegs_field *vp,
                      mlegs_field *vz) {
  return tp2vec_impl(psi, chi, vr, vp, vz);
}
int mlegs_b200_tp2curlvec(const mlegs_field *psi, const mlegs_field *chi, mlegs_field *wr, mlegs_field *wp,
                          mlegs_field *wz) {
  return tp2curlvec_impl(psi, chi, wr, wp, wz);
}

int mlegs_b200_axpby(mlegs_field *y, double a, const mlegs_field *x, double b) {
  MLEGS_TRY(ready());
  return lin(0, y, x, nullptr, nullptr, a, b, 0.0);
}

int mlegs_b200_is_finite(const mlegs_field *s, int *all_finite) {
  MLEGS_TRY(ready());
  Context &c = ctx();
  CUDA_TRY(cudaMemsetAsync(c.d_flag, 0, sizeof(int), strm()));
  MLEGS_TRY(launch_finite((const cplx *)s->e, nelem(s), c.d_flag, strm()));
  int h[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(h, c.d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, strm()));
  CUDA_TRY(cudaStreamSynchronize(strm()));
  *all_finite = h[0] ? 0 : 1;
  if (h[1]) {
    CUDA_TRY(cudaMemsetAsync(c.d_flag + 1, 0, sizeof(int), strm()));
    return fail(MLEGS_E_ARG, "lurc: lu factorization resulted in failure");
  }
  return MLEGS_OK;
}

}  // extern "C"
