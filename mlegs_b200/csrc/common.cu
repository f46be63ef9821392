// Library state: error reporting, the device-resident transform kit, scalar storage.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "kernels.h"

namespace mlegs {

static thread_local std::string g_error;
long long g_launches = 0;

void set_error(const std::string &msg) { g_error = msg; }
int fail(int code, const std::string &msg) {
  g_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  g_error = buf;
  cudaGetLastError();
  return MLEGS_E_CUDA;
}

Context &ctx() {
  static Context c;
  return c;
}

static cudaStream_t stream() { return (cudaStream_t)ctx().stream; }

int chop_index(const mlegs_field *s, ChopIdx *ci) {
  Context &c = ctx();
  ci->nrc = c.p.nrchop + s->nrchop_offset;
  ci->npc = c.p.npchop + s->npchop_offset;
  ci->nzc = c.chopzl + s->nzchop_offset;
  ci->nzcu = c.chopzu - s->nzchop_offset;
  if (ci->nrc > c.nrdim) return fail(MLEGS_E_ARG, "chop_index: chopping in r too large");
  if (ci->npc > c.npdim) return fail(MLEGS_E_ARG, "chop_index: chopping in p too large");
  if (2 * s->nzchop_offset > ci->nzcu - ci->nzc) return fail(MLEGS_E_ARG, "chop_index: chopping in z too large");
  return MLEGS_OK;
}

template <typename T>
static int upload(T **dst, const T *src, size_t n) {
  CUDA_TRY(cudaMalloc((void **)dst, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) CUDA_TRY(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return MLEGS_OK;
}

static int upload_twiddles(double **dst, int n) {
  std::vector<double> tw(2 * (size_t)std::max(n, 1));
  const long double twopi = 2.0L * acosl(-1.0L);
  for (int j = 0; j < n; ++j) {
    long double ang = twopi * (long double)j / (long double)n;
    tw[2 * j] = (double)cosl(ang);
    tw[2 * j + 1] = (double)(-sinl(ang));
  }
  return upload(dst, tw.data(), tw.size());
}

static bool smooth235(int n) {
  while (n % 2 == 0 || n % 3 == 0 || n % 5 == 0) {
    if (n % 2 == 0) n /= 2;
    if (n % 3 == 0) n /= 3;
    if (n % 5 == 0) n /= 5;
  }
  return n == 1;
}

// tfm_kit_init's argument checks, /root/reference/src/submodules/mlegs_spectfm_init.f90:35-62
int validate_params(const mlegs_params *p) {
  if (!(p->nr > 0 && p->nr % 2 == 0)) return fail(MLEGS_E_ARG, "tfm_kit_init: nr must be even");
  if (!(p->np > 0 && (p->np == 1 || p->np % 2 == 0))) return fail(MLEGS_E_ARG, "tfm_kit_init: np must be even");
  if (p->np != 1 && !smooth235(p->np))
    return fail(MLEGS_E_ARG, "tfm_kit_init: np must only have factors of 2, 3 and 5");
  if (!(p->nz > 0 && (p->nz == 1 || p->nz % 2 == 0))) return fail(MLEGS_E_ARG, "tfm_kit_init: nz must be even");
  if (p->nz != 1 && !smooth235(p->nz))
    return fail(MLEGS_E_ARG, "tfm_kit_init: nz must only have factors of 2, 3 and 5");
  if (p->nrchop > p->nr) return fail(MLEGS_E_ARG, "tfm_kit_init: nrchop must be smaller than or equal to nr");
  if (p->np != 1 && p->npchop * 2 > p->np + 2)
    return fail(MLEGS_E_ARG, "tfm_kit_init: npchop <= np/2 + 1 must be satisfied");
  if (p->nz != 1 && p->nzchop * 2 > p->nz + 2)
    return fail(MLEGS_E_ARG, "tfm_kit_init: nzchop <= nz/2 + 1 must be satisfied");
  if (p->nrchop <= 0 || p->npchop <= 0 || p->nzchop <= 0)
    return fail(MLEGS_E_ARG, "read_input: invalid chopping sizes");
  if (p->ell <= 0.0 || p->zlen <= 0.0) return fail(MLEGS_E_ARG, "read_input: ell and zlen must be positive");
  if (p->hyperpow != 0 && p->hyperpow != 4 && p->hyperpow != 6 && p->hyperpow != 8)
    return fail(MLEGS_E_ARG, "read_input: hyperpow must be zero, 4, 6, or 8");
  return MLEGS_OK;
}

int build_operator_tables();   // banded.cu
int leg_ws_build_pfw();         // legendre_ws.cu
void leg_ws_reset();

static int free_all() {
  Context &c = ctx();
  band_solve_cache_clear();
  leg_ws_reset();
  void *ptrs[] = {c.d_x, c.d_w, c.d_lnx, c.d_r, c.d_lognorm, c.d_pf, c.d_pfw, c.d_at0, c.d_at1, c.d_ak, c.d_tw_p,
                  c.d_tw_z, c.d_del2h, c.d_xxdx, c.d_vtab, c.d_dtab, c.d_scratch[0], c.d_scratch[1],
                  c.d_scratch[2], c.d_scratch[3], c.d_scratch[4], c.d_scratch[5], c.d_red, c.d_solve_ws, c.d_flags,
                  c.d_flag, c.d_colstart, c.d_cossin_p};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  for (void *p : c.d_batch)
    if (p) cudaFree(p);
  if (c.h_red) cudaFreeHost(c.h_red);
  void *st = c.stream;
  c = Context();
  c.stream = st;
  return MLEGS_OK;
}

}  // namespace mlegs

namespace mlegs {
int field_mstride(const mlegs_field *f) {
  Context &c = ctx();
  return (c.nranks > 1 && f->loc_sz[1] != c.npdim) ? c.nranks : 1;
}
void field_set_layout(mlegs_field *f, bool physical) {
  Context &c = ctx();
  f->glb_sz[0] = c.nrdim;
  f->glb_sz[1] = c.npdim;
  f->glb_sz[2] = c.nzdim;
  if (physical) {   // axis_comm (1,0,2): r distributed (apps/vortical_flow_3d.f90:81)
    f->loc_sz[0] = c.r_cnt[c.rank];
    f->loc_sz[1] = c.npdim;
    f->loc_sz[2] = c.nzdim;
    f->loc_st[0] = c.r_off[c.rank];
    f->loc_st[1] = 0;
    f->loc_st[2] = 0;
    f->axis_comm[0] = 1;
    f->axis_comm[1] = 0;
    f->axis_comm[2] = 2;
  } else {          // axis_comm (2,1,0): m distributed (apps/vortical_flow_3d.f90:84)
    f->loc_sz[0] = c.nrdim;
    f->loc_sz[1] = c.m_cnt[c.rank];
    f->loc_sz[2] = c.nzdim;
    f->loc_st[0] = 0;
    f->loc_st[1] = c.nranks > 1 ? c.rank : 0;   // first column; the following ones are nranks apart
    f->loc_st[2] = 0;
    f->axis_comm[0] = 2;
    f->axis_comm[1] = 1;
    f->axis_comm[2] = 0;
  }
}
}  // namespace mlegs

using namespace mlegs;

extern "C" {

const char *mlegs_b200_last_error(void) { return g_error.c_str(); }
int mlegs_b200_version(void) { return 100; }

int mlegs_b200_set_stream(void *cuda_stream) {
  ctx().stream = cuda_stream;
  return MLEGS_OK;
}

int mlegs_b200_device_sync(void) {
  CUDA_TRY(cudaStreamSynchronize(stream()));
  return MLEGS_OK;
}

long long mlegs_b200_launch_count(int reset) {
  long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int mlegs_b200_tfm_tables(const mlegs_params *p, double *x, double *w, double *ln, double *r, double *lognorm,
                          double *pf, double *at0, double *at1, double *ak) {
  MLEGS_TRY(validate_params(p));
  return build_tfm_tables(p, x, w, ln, r, lognorm, pf, at0, at1, ak);
}

int mlegs_b200_tfm_tables_cached(const mlegs_params *p, const char *cache_dir, double *x, double *w, double *ln,
                                 double *r, double *lognorm, double *pf, double *at0, double *at1, double *ak,
                                 int *from_cache) {
  MLEGS_TRY(validate_params(p));
  return build_tfm_tables_cached(p, cache_dir, x, w, ln, r, lognorm, pf, at0, at1, ak, from_cache);
}

int mlegs_b200_init(const mlegs_params *p, const double *x, const double *w, const double *lognorm,
                    const double *pf, const double *at0, const double *at1, int rank, int nranks) {
  MLEGS_TRY(validate_params(p));
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(MLEGS_E_CUDA, "mlegs_b200_init: no CUDA device (this library has no CPU fallback)");
  }
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(MLEGS_E_COMM, "mlegs_b200_init: bad rank/nranks");
  free_all();
  Context &c = ctx();
  c.p = *p;
  c.rank = rank;
  c.nranks = nranks;
  c.nrdim = p->nr + std::max(3, p->hyperpow);   // sinit:117
  c.npdim = p->np / 2 + 1;
  c.nzdim = p->nz;
  c.nrh = p->nr / 2;
  c.ne = p->nrchop + 14;
  c.chopzl = p->nzchop;                         // sinit:138-139
  c.chopzu = p->nz - p->nzchop + 2;
  const int nr = p->nr;
  c.h_x.assign(x, x + nr);
  c.h_w.assign(w, w + nr);
  c.h_ln.resize(nr);
  std::vector<double> h_r(nr);
  for (int i = 0; i < nr; ++i) {
    c.h_ln[i] = -std::log(1.0 - x[i]);                               // sinit:68
    h_r[i] = p->ell * std::sqrt((1.0 + x[i]) / (1.0 - x[i]));        // sinit:112
  }
  c.h_lognorm.assign(lognorm, lognorm + (size_t)c.ne * p->npchop);
  c.h_at0.assign(at0, at0 + p->nrchop);
  c.h_at1.assign(at1, at1 + p->nrchop);
  const double pi = std::acos(-1.0);
  c.h_ak.resize(p->nz);
  for (int i = 0; i < p->nz; ++i) c.h_ak[i] = 2.0 * pi / p->zlen * (double)(i - p->nz);        // sinit:133
  for (int i = 0; i <= p->nz / 2 && i < p->nz; ++i) c.h_ak[i] = 2.0 * pi / p->zlen * (double)i;  // sinit:134

  MLEGS_TRY(upload(&c.d_x, c.h_x.data(), nr));
  MLEGS_TRY(upload(&c.d_w, c.h_w.data(), nr));
  MLEGS_TRY(upload(&c.d_lnx, c.h_ln.data(), nr));
  MLEGS_TRY(upload(&c.d_r, h_r.data(), nr));
  MLEGS_TRY(upload(&c.d_lognorm, c.h_lognorm.data(), c.h_lognorm.size()));
  MLEGS_TRY(upload(&c.d_pf, pf, (size_t)c.nrh * c.ne * p->npchop));
  MLEGS_TRY(upload(&c.d_at0, c.h_at0.data(), c.h_at0.size()));
  MLEGS_TRY(upload(&c.d_at1, c.h_at1.data(), c.h_at1.size()));
  MLEGS_TRY(upload(&c.d_ak, c.h_ak.data(), c.h_ak.size()));
  MLEGS_TRY(upload_twiddles(&c.d_tw_p, p->np));
  MLEGS_TRY(upload_twiddles(&c.d_tw_z, p->nz));
  if (p->np > 1) MLEGS_TRY(make_fft_plan(p->np / 2, 1, &c.plan_p));
  if (p->nz > 1) MLEGS_TRY(make_fft_plan(p->nz, 0, &c.plan_z));
  MLEGS_TRY(setup_fft_kernels());
  MLEGS_TRY(setup_leg_kernels());
  MLEGS_TRY(leg_ws_build_pfw());

  c.r_cnt.resize(nranks);
  c.r_off.resize(nranks);
  c.m_cnt.resize(nranks);
  c.m_off.resize(nranks);
  for (int q = 0; q < nranks; ++q) {
    decompose(c.nrdim, nranks, q, &c.r_cnt[q], &c.r_off[q]);
    c.m_cnt[q] = q < c.npdim ? (c.npdim - q + nranks - 1) / nranks : 0;   // m = q, q + P, ...
    c.m_off[q] = q == 0 ? 0 : c.m_off[q - 1] + c.m_cnt[q - 1];
  }
  // field-sized scratch: the larger of the two slab shapes
  size_t n_ppp = (size_t)c.r_cnt[0] * c.npdim * c.nzdim;
  size_t n_fff = (size_t)c.nrdim * c.m_cnt[0] * c.nzdim;
  c.field_bytes = std::max(n_ppp, n_fff) * sizeof(cplx);
  for (int i = 0; i < 6; ++i) CUDA_TRY(cudaMalloc(&c.d_scratch[i], c.field_bytes));
  CUDA_TRY(cudaMalloc((void **)&c.d_red, Context::RED_DOUBLES * sizeof(double)));
  CUDA_TRY(cudaMallocHost((void **)&c.h_red, Context::RED_DOUBLES * sizeof(double)));
  c.ready = true;
  MLEGS_TRY(build_operator_tables());
  return MLEGS_OK;
}

int mlegs_b200_finalize(void) { return free_all(); }

int mlegs_b200_update_params(const mlegs_params *p) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  c.p.visc = p->visc;
  c.p.hypervisc = p->hypervisc;
  c.p.is_svv = p->is_svv;
  c.p.svv_cutoff = p->svv_cutoff;
  c.p.svv_target = p->svv_target;
  c.p.svv_strength = p->svv_strength;
  c.p.svv_relax = p->svv_relax;
  return MLEGS_OK;
}

// ---- scalar storage ------------------------------------------------------------------------

static size_t field_elems(const mlegs_field *f) { return (size_t)f->loc_sz[0] * f->loc_sz[1] * f->loc_sz[2]; }


static bool g_managed = false;

int mlegs_b200_use_managed(int on) {
  g_managed = on != 0;
  return MLEGS_OK;
}

static int alloc_field_bytes(void **p) {
  Context &c = ctx();
  if (!g_managed) {
    CUDA_TRY(cudaMalloc(p, c.field_bytes));
    return MLEGS_OK;
  }
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaMallocManaged(p, c.field_bytes, cudaMemAttachGlobal));
  CUDA_TRY(cudaMemAdvise(*p, c.field_bytes, cudaMemAdviseSetPreferredLocation, dev));
  CUDA_TRY(cudaMemPrefetchAsync(*p, c.field_bytes, dev, stream()));
  return MLEGS_OK;
}

int mlegs_b200_field_alloc(mlegs_field *f, const char *space3) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  memset(f, 0, sizeof(*f));
  bool physical = strncmp(space3, "PPP", 3) == 0;
  field_set_layout(f, physical);
  // every field buffer has the size of the larger slab shape so an exchange can reuse it
  MLEGS_TRY(alloc_field_bytes(&f->e));
  CUDA_TRY(cudaMemsetAsync(f->e, 0, c.field_bytes, stream()));
  f->ln = 0.0;
  memcpy(f->space, space3, 3);
  f->space[3] = 0;
  return MLEGS_OK;
}

int mlegs_b200_field_free(mlegs_field *f) {
  if (f->e) CUDA_TRY(cudaFree(f->e));
  f->e = nullptr;
  f->ln = 0.0;
  f->nrchop_offset = f->npchop_offset = f->nzchop_offset = 0;
  f->space[0] = 0;
  return MLEGS_OK;
}

int mlegs_b200_field_copy(mlegs_field *dst, const mlegs_field *src) {
  if (!src->e) return MLEGS_OK;   // scalar_copy warns and returns, mlegs_scalar_init.f90:107-111
  if (!dst->e) {
    void *keep = nullptr;
    MLEGS_TRY(alloc_field_bytes(&keep));
    dst->e = keep;
  }
  void *e = dst->e;
  *dst = *src;
  dst->e = e;
  CUDA_TRY(cudaMemcpyAsync(dst->e, src->e, field_elems(src) * sizeof(cplx), cudaMemcpyDeviceToDevice, stream()));
  return MLEGS_OK;
}

int mlegs_b200_field_zero(mlegs_field *f) {
  CUDA_TRY(cudaMemsetAsync(f->e, 0, field_elems(f) * sizeof(cplx), stream()));
  return MLEGS_OK;
}

int mlegs_b200_field_upload(mlegs_field *f, const void *host_e) {
  CUDA_TRY(cudaMemcpyAsync(f->e, host_e, field_elems(f) * sizeof(cplx), cudaMemcpyHostToDevice, stream()));
  CUDA_TRY(cudaStreamSynchronize(stream()));
  return MLEGS_OK;
}

int mlegs_b200_field_download(const mlegs_field *f, void *host_e) {
  CUDA_TRY(cudaMemcpyAsync(host_e, f->e, field_elems(f) * sizeof(cplx), cudaMemcpyDeviceToHost, stream()));
  CUDA_TRY(cudaStreamSynchronize(stream()));
  return MLEGS_OK;
}

int mlegs_b200_field_chop_offset(mlegs_field *f, int iof1, int iof2, int iof3) {
  f->nrchop_offset = iof1;   // mlegs_scalar_init.f90:84-103
  f->npchop_offset = iof2;
  f->nzchop_offset = iof3;
  return MLEGS_OK;
}

int mlegs_b200_host_register(void *host_ptr, size_t bytes) {
  CUDA_TRY(cudaHostRegister(host_ptr, bytes, cudaHostRegisterDefault));
  return MLEGS_OK;
}
int mlegs_b200_host_unregister(void *host_ptr) {
  CUDA_TRY(cudaHostUnregister(host_ptr));
  return MLEGS_OK;
}

}  // extern "C"
