// trans(): the PPP <-> PFP <-> FFP <-> FFF state machine of
// /root/reference/src/submodules/mlegs_scalar_ops.f90:157-235, composed from the FFT and Legendre
// kernels.  On one rank the two scalar_exchange calls are pure re-labellings (the block already is
// the whole (nrdim, npdim, nzdim) array); on several ranks the (2,1)/(1,2) exchange is the
// all-to-all in dist.cu and the (1,3)/(3,1) exchange is a no-op in the slab layout.
#include <algorithm>
#include <cstring>
#include <vector>

#include "dist_dev.cuh"

namespace mlegs {

int exchange_slab(mlegs_field *s, int axis_old, int axis_new, const void *src, void **landed);   // dist.cu
int dist_begin_put(PeerTable *t, void **landed);                                                  // dist.cu

static int space_id(const char *sp) {
  if (!strncmp(sp, "PPP", 3)) return 0;
  if (!strncmp(sp, "PFP", 3)) return 1;
  if (!strncmp(sp, "FFP", 3)) return 2;
  if (!strncmp(sp, "FFF", 3)) return 3;
  return -1;
}
static const char *kSpaceName[4] = {"PPP", "PFP", "FFP", "FFF"};

// space label + the axis_comm labels the reference's trans leaves behind (ops:185-233 with the exchanges of
// dist:52-58): PPP (1,0,2), PFP/FFP (0,1,2), FFF (2,1,0)
static void set_space(mlegs_field *s, int id) {
  memcpy(s->space, kSpaceName[id], 3);
  s->space[3] = 0;
  static const int labels[4][3] = {{1, 0, 2}, {0, 1, 2}, {0, 1, 2}, {2, 1, 0}};
  for (int a = 0; a < 3; ++a) s->axis_comm[a] = labels[id][a];
}

// the azimuthal c2r FFT that follows a fused exchange(1,2) reads the window's transit layout (columns grouped by the
// rank that owns them, dist_dev.cuh)
static void set_transit_perm(RowScale *rs) {
  Context &c = ctx();
  rs->perm_p = c.nranks;
  for (int q = 0; q < c.nranks && q < 16; ++q) rs->perm_off[q] = c.m_off[q];
}

static int rtrans_args(const mlegs_field *s, const char *who, LegArgs *a) {
  Context &c = ctx();
  int nrc = c.p.nrchop + s->nrchop_offset;
  int npc = c.p.npchop + s->npchop_offset;
  if (nrc > c.nrdim) return fail(MLEGS_E_ARG, std::string(who) + ": chopping in r too large");
  if (npc > c.npdim) return fail(MLEGS_E_ARG, std::string(who) + ": chopping in p too large");
  a->pf = c.d_pf;
  a->w = c.d_w;
  a->lnx = c.d_lnx;
  a->nr = c.p.nr;
  a->nrh = c.nrh;
  a->ne = c.ne;
  a->nrl = s->loc_sz[0];
  a->npl = s->loc_sz[1];
  a->m0 = s->loc_st[1];
  a->ms = field_mstride(s);
  a->nzl = s->loc_sz[2];
  a->nrc = nrc;
  a->npc = npc;
  a->nrdim = c.nrdim;
  a->lnval = s->ln;
  a->swap_parity = 0;
  a->skip_m0 = 0;
  a->peer = nullptr;
  a->npdim = c.npdim;
  return MLEGS_OK;
}

// One stage, reading `src` and writing `dst` (may be equal for the FFT stages).  fb != nullptr: the stage runs on
// the fb->n scalars of the batch in one launch (all with the layout/metadata of `s`; src/dst are ignored).
int stage_phi(const mlegs_field *s, bool forward, const cplx *src, cplx *dst, const FieldBatch *fb = nullptr,
              const RowScale *rs = nullptr);
int stage_phi(const mlegs_field *s, bool forward, const cplx *src, cplx *dst, const FieldBatch *fb, const RowScale *rs) {
  Context &c = ctx();
  cudaStream_t st = (cudaStream_t)c.stream;
  long long rows = s->loc_sz[0];
  return launch_fft_lines(forward ? FFT_R2C_FWD : FFT_C2R_BWD, c.plan_p, src, dst, rows, rows, s->loc_sz[2],
                          rows * (long long)s->loc_sz[1], c.d_tw_p, c.p.np, forward ? 1.0 / c.p.np : 1.0, st, fb, rs);
}

int stage_z(const mlegs_field *s, bool forward, const cplx *src, cplx *dst, const FieldBatch *fb = nullptr);
int stage_z(const mlegs_field *s, bool forward, const cplx *src, cplx *dst, const FieldBatch *fb) {
  Context &c = ctx();
  cudaStream_t st = (cudaStream_t)c.stream;
  long long plane = (long long)s->loc_sz[0] * s->loc_sz[1];
  return launch_fft_lines(forward ? FFT_C2C_FWD : FFT_C2C_BWD, c.plan_z, src, dst, plane, plane, 1, 0, c.d_tw_z,
                          c.p.nz, forward ? 1.0 / c.p.nz : 1.0, st, fb);
}

// Axial FFT restricted to the retained lines (rows < nn(m) of every local column m < npc).  Valid when the
// other lines are known zeros that stay in place (forward, right after rtrans_forward) or are never read
// again (backward, right before rtrans_backward).
static int stage_z_compact(const mlegs_field *s, bool forward, const cplx *src, cplx *dst,
                           const FieldBatch *fb = nullptr) {
  Context &c = ctx();
  cudaStream_t st = (cudaStream_t)c.stream;
  const int nrc = c.p.nrchop + s->nrchop_offset, npc = c.p.npchop + s->npchop_offset;
  const int npl = s->loc_sz[1], m0 = s->loc_st[1], ms = field_mstride(s);
  if (c.cs_nrc != nrc || c.cs_npc != npc || c.cs_ncols != npl || !c.d_colstart) {
    std::vector<int> h(npl + 1, 0);
    for (int j = 0; j < npl; ++j) {
      int m = m0 + j * ms;
      int nn = (m < npc) ? std::max(std::min(nrc, nrc - m), 0) : 0;
      nn = std::min(nn, s->loc_sz[0]);
      h[j + 1] = h[j] + nn;
    }
    if (c.d_colstart) CUDA_TRY(cudaFree(c.d_colstart));
    c.d_colstart = nullptr;
    CUDA_TRY(cudaMalloc((void **)&c.d_colstart, (npl + 1) * sizeof(int)));
    CUDA_TRY(cudaMemcpyAsync(c.d_colstart, h.data(), (npl + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));   // h goes out of scope
    c.cs_nrc = nrc;
    c.cs_npc = npc;
    c.cs_ncols = npl;
    c.cs_total = h[npl];
  }
  long long plane = (long long)s->loc_sz[0] * s->loc_sz[1];
  return launch_fft_z_compact(forward ? FFT_C2C_FWD : FFT_C2C_BWD, c.plan_z, src, dst, c.d_colstart, npl, s->loc_sz[0],
                              c.cs_total, plane, c.d_tw_z, c.p.nz, forward ? 1.0 / c.p.nz : 1.0, st, fb);
}

int stage_r(const mlegs_field *s, bool forward, const cplx *src, cplx *dst, const PeerTable *peer = nullptr,
            const FieldBatch *fb = nullptr);
int stage_r(const mlegs_field *s, bool forward, const cplx *src, cplx *dst, const PeerTable *peer,
            const FieldBatch *fb) {
  LegArgs a;
  MLEGS_TRY(rtrans_args(s, forward ? "rtrans_forward" : "rtrans_backward", &a));
  a.in = src;
  a.out = dst;
  a.peer = peer;
  if (fb) a.fb = *fb;
  cudaStream_t st = (cudaStream_t)ctx().stream;
  return forward ? launch_leg_forward(a, st) : launch_leg_backward(a, st);
}

int trans_impl(mlegs_field *s, const char *to) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  int cur = space_id(s->space), dst = space_id(to);
  if (cur < 0)
    return fail(MLEGS_E_ARG, "trans: scalar space info corrupted (only accepting PPP, PFP, FFP and FFF)");
  if (dst < 0) return fail(MLEGS_E_ARG, "trans: only taking PPP, PFP, FFP and FFF for spectral transformation");
  if (cur == dst) return MLEGS_OK;
  cudaStream_t st = (cudaStream_t)c.stream;
  const bool has_p = c.p.np > 1, has_z = c.p.nz > 1;
  const bool multi = c.nranks > 1 && has_p;

  const bool compact_ok = has_z && fft_reg_supported(c.plan_z.n);
  bool rows_zero = false, exchanged = false;
  cplx *home = (cplx *)s->e;
  cplx *tmp = (cplx *)c.d_scratch[0];
  cplx *at = home;   // where the data currently lives
  auto other = [&](cplx *p) { return p == home ? tmp : home; };

  if (cur < dst) {   // forward, ops:185-208
    while (cur < dst) {
      if (cur == 0) {
        if (has_p) {
          if (multi && fft_reg_supported(c.plan_p.n)) {
            // the FFT's stores ARE the (2,1) exchange: every output column goes straight into the window of the
            // rank that owns that m, and the kernel ends with the exchange barrier
            PeerTable pt;
            void *landed = nullptr;
            MLEGS_TRY(dist_begin_put(&pt, &landed));
            const long long rows = s->loc_sz[0];
            MLEGS_TRY(launch_fft_phi_forward_put(c.plan_p, at, rows, s->loc_sz[2], rows * (long long)s->loc_sz[1],
                                                 c.d_tw_p, c.p.np, 1.0 / c.p.np, pt, c.nrdim, st));
            field_set_layout(s, false);
            at = (cplx *)landed;
          } else if (multi) {
            // FFT in place, then the (2,1) exchange moves the block into this rank's receive window
            MLEGS_TRY(stage_phi(s, true, at, at));
            void *landed = nullptr;
            MLEGS_TRY(exchange_slab(s, 2, 1, at, &landed));
            at = (cplx *)landed;
          } else {
            // go out of place when the Legendre stage follows, so that it lands back at home and the axial
            // stage can then run in place on the retained lines only
            cplx *o = (dst >= 2) ? other(at) : at;
            MLEGS_TRY(stage_phi(s, true, at, o));
            at = o;
          }
        }
      } else if (cur == 1) {
        MLEGS_TRY(stage_r(s, true, at, other(at)));   // ln removal (ops:193-195) is fused into the load
        at = other(at);
        rows_zero = true;                             // rows >= nn(m) are exact zeros now
      } else if (cur == 2) {
        if (has_z) {
          cplx *o = (at == home) ? at : home;   // land at home whenever possible
          if (rows_zero && o == at && compact_ok)
            MLEGS_TRY(stage_z_compact(s, true, at, o));
          else
            MLEGS_TRY(stage_z(s, true, at, o));
          at = o;
        }
      }
      ++cur;
      set_space(s, cur);
    }
  } else {           // backward, ops:210-233
    while (cur > dst) {
      if (cur == 3) {
        if (has_z) {
          // if the Legendre stage follows, go out of place so that it lands back at home
          cplx *o = (dst <= 1) ? other(at) : at;
          if (dst <= 1 && compact_ok)
            MLEGS_TRY(stage_z_compact(s, false, at, o));   // rtrans_backward reads the retained rows only
          else
            MLEGS_TRY(stage_z(s, false, at, o));
          at = o;
        }
      } else if (cur == 2) {
        if (multi && dst == 0 && fft_reg_supported(c.plan_p.n)) {
          // the epilogue's stores ARE the (1,2) exchange (rows go to the ranks that own them in physical space)
          PeerTable pt;
          void *landed = nullptr;
          MLEGS_TRY(dist_begin_put(&pt, &landed));
          MLEGS_TRY(stage_r(s, false, at, other(at), &pt));   // other(at): staging buffer of the (1,2) exchange
          field_set_layout(s, true);
          at = (cplx *)landed;
          exchanged = true;
        } else {
          MLEGS_TRY(stage_r(s, false, at, other(at)));  // + ln term (ops:219-221) fused into the epilogue
          at = other(at);
        }
      } else if (cur == 1) {
        if (has_p) {
          if (multi) {
            if (!exchanged) {
              void *landed = nullptr;
              MLEGS_TRY(exchange_slab(s, 1, 2, at, &landed));
              at = (cplx *)landed;
            }
            cplx *o = home;
            RowScale rs;
            if (exchanged) set_transit_perm(&rs);   // the window holds the transit layout of the fused exchange
            MLEGS_TRY(stage_phi(s, false, at, o, nullptr, exchanged ? &rs : nullptr));
            at = o;
          } else {
            cplx *o = home;
            MLEGS_TRY(stage_phi(s, false, at, o));
            at = o;
          }
        }
      }
      --cur;
      set_space(s, cur);
    }
  }
  if (at != home) {
    size_t n = (size_t)s->loc_sz[0] * s->loc_sz[1] * s->loc_sz[2];
    CUDA_TRY(cudaMemcpyAsync(home, at, n * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
  }
  return MLEGS_OK;
}

// ------------------------------------------------------------------------------------------------
// trans() of several scalars at once (one rank): the same state machine, but every stage is ONE launch
// over all scalars of the batch (field index = a grid dimension).  The scalars must agree in space,
// chopping offsets and layout -- e.g. the components of a vector field (ops:1503-1505) or the fields of
// a multi-scalar app.  Arithmetic per scalar is identical to trans_impl (bit-identical results).
// ------------------------------------------------------------------------------------------------
static int batch_scratch(int n, cplx **tmp) {
  Context &c = ctx();
  static_assert(MLEGS_MAXB <= sizeof(c.d_batch) / sizeof(c.d_batch[0]), "batch scratch");
  for (int i = 0; i < n; ++i) {   // field-sized, allocated on first use, freed with the context
    if (!c.d_batch[i]) CUDA_TRY(cudaMalloc(&c.d_batch[i], c.field_bytes));
    tmp[i] = (cplx *)c.d_batch[i];
  }
  return MLEGS_OK;
}

int dist_window_batch();       // dist.cu
size_t dist_field_stride();

// One group of at most MLEGS_MAXB scalars (several ranks: at most dist_window_batch()).  Every scalar has a current
// location (its own buffer, its batch scratch buffer, or -- right after a fused exchange -- its slab of this rank's
// receive window); stages go out of place whenever that lets the last one land at home.
// Options of a batched transform used by the vector operations: scalars whose rows are multiplied by r on the way
// into the forward azimuthal FFT / divided by r on the way out of the backward one (bit i of rs_mask), and scalars
// whose input is read from another (read-only) buffer than their own (src[i] != nullptr: the first stage then goes out
// of place, which replaces a field copy).
struct TransOpts {
  unsigned rs_mask = 0;
  const void *src[MLEGS_MAXB] = {};
};

static int trans_group(int n, mlegs_field *const *s, int cur, int dst, const TransOpts *opt = nullptr, int i0 = 0) {
  Context &c = ctx();
  cudaStream_t st = (cudaStream_t)c.stream;
  const bool has_p = c.p.np > 1, has_z = c.p.nz > 1;
  const bool multi = c.nranks > 1 && has_p;
  const bool compact_ok = has_z && fft_reg_supported(c.plan_z.n);
  cplx *home[MLEGS_MAXB], *tmp[MLEGS_MAXB], *at[MLEGS_MAXB];
  MLEGS_TRY(batch_scratch(n, tmp));
  for (int i = 0; i < n; ++i) {
    home[i] = (cplx *)s[i]->e;
    at[i] = (opt && opt->src[i0 + i]) ? (cplx *)const_cast<void *>(opt->src[i0 + i]) : home[i];   // never written to
  }
  const bool foreign_src = at[0] != home[0];
  RowScale rs;
  const unsigned nmask = n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
  if (opt && ((opt->rs_mask >> i0) & nmask)) {
    rs.r = c.d_r;
    rs.mask = (opt->rs_mask >> i0) & nmask;
    rs.nr = c.p.nr;
  }
  bool rows_zero = false, in_transit = false;
  const mlegs_field *s0 = s[0];
  auto other = [&](int i) { return at[i] == home[i] ? tmp[i] : home[i]; };
  // batch reading from where the data is and writing in place, or to the buffer the data is not in
  auto make = [&](bool inplace, FieldBatch *fb) {
    fb->n = n;
    for (int i = 0; i < n; ++i) {
      fb->in[i] = at[i];
      fb->out[i] = inplace ? at[i] : other(i);
      fb->ln[i] = s[i]->ln;
    }
  };
  auto moved = [&](const FieldBatch &fb) {
    for (int i = 0; i < n; ++i) at[i] = fb.out[i];
  };
  // the scalars now live in consecutive slabs of this rank's receive window, in the other slab layout
  auto landed_in_window = [&](void *landed, bool physical) {
    for (int i = 0; i < n; ++i) {
      at[i] = reinterpret_cast<cplx *>(reinterpret_cast<char *>(landed) + (size_t)i * dist_field_stride());
      field_set_layout(s[i], physical);
    }
  };
  FieldBatch fb;
  if (cur < dst) {   // forward, ops:185-208
    while (cur < dst) {
      if (cur == 0 && has_p) {
        if (multi) {
          // the FFT's stores ARE the (2,1) exchange of all n scalars: one launch, one barrier
          PeerTable pt;
          void *landed = nullptr;
          MLEGS_TRY(dist_begin_put(&pt, &landed));
          make(true, &fb);
          const long long rows = s0->loc_sz[0];
          rs.mode = rs.mask ? 1 : 0;
          rs.r0 = s0->loc_st[0];
          MLEGS_TRY(launch_fft_phi_forward_put(c.plan_p, nullptr, rows, s0->loc_sz[2], rows * (long long)s0->loc_sz[1],
                                               c.d_tw_p, c.p.np, 1.0 / c.p.np, pt, c.nrdim, st, &fb, &rs));
          landed_in_window(landed, false);
        } else {
          const bool inplace = !(dst >= 2) && !foreign_src;
          make(inplace, &fb);
          rs.mode = rs.mask ? 1 : 0;
          rs.r0 = s0->loc_st[0];
          MLEGS_TRY(stage_phi(s0, true, nullptr, nullptr, &fb, &rs));
          moved(fb);
        }
      } else if (cur == 1) {
        make(false, &fb);
        MLEGS_TRY(stage_r(s0, true, nullptr, nullptr, nullptr, &fb));
        moved(fb);
        rows_zero = true;
      } else if (cur == 2 && has_z) {
        const bool inplace = at[0] == home[0];   // land at home whenever possible
        make(inplace, &fb);
        if (rows_zero && inplace && compact_ok)
          MLEGS_TRY(stage_z_compact(s0, true, nullptr, nullptr, &fb));
        else
          MLEGS_TRY(stage_z(s0, true, nullptr, nullptr, &fb));
        moved(fb);
      }
      ++cur;
    }
  } else {           // backward, ops:210-233
    while (cur > dst) {
      if (cur == 3 && has_z) {
        const bool inplace = !(dst <= 1);
        make(inplace, &fb);
        if (dst <= 1 && compact_ok)
          MLEGS_TRY(stage_z_compact(s0, false, nullptr, nullptr, &fb));
        else
          MLEGS_TRY(stage_z(s0, false, nullptr, nullptr, &fb));
        moved(fb);
      } else if (cur == 2) {
        if (multi && dst == 0) {
          // the epilogue's stores ARE the (1,2) exchange (rows go to the ranks that own them in physical space)
          PeerTable pt;
          void *landed = nullptr;
          MLEGS_TRY(dist_begin_put(&pt, &landed));
          make(false, &fb);   // the other buffer of every scalar stages its rows in destination order
          MLEGS_TRY(stage_r(s0, false, nullptr, nullptr, &pt, &fb));
          landed_in_window(landed, true);
          in_transit = true;                // columns grouped by owner: the azimuthal FFT below reads them permuted
        } else {
          make(false, &fb);
          MLEGS_TRY(stage_r(s0, false, nullptr, nullptr, nullptr, &fb));
          moved(fb);
        }
      } else if (cur == 1 && has_p) {
        const bool inplace = at[0] == home[0];
        make(inplace, &fb);
        if (!inplace)
          for (int i = 0; i < n; ++i) fb.out[i] = home[i];
        rs.mode = rs.mask ? 2 : 0;
        rs.r0 = s[0]->loc_st[0];
        if (in_transit) set_transit_perm(&rs);
        MLEGS_TRY(stage_phi(s[0], false, nullptr, nullptr, &fb, &rs));
        moved(fb);
      }
      --cur;
    }
  }
  for (int i = 0; i < n; ++i) {
    set_space(s[i], dst);
    if (at[i] != home[i]) {
      size_t ne = (size_t)s[i]->loc_sz[0] * s[i]->loc_sz[1] * s[i]->loc_sz[2];
      CUDA_TRY(cudaMemcpyAsync(home[i], at[i], ne * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
    }
  }
  return MLEGS_OK;
}

static int trans_many_opt(int n, mlegs_field *const *s, const char *to, const TransOpts *opt);
int trans_many_impl(int n, mlegs_field *const *s, const char *to) { return trans_many_opt(n, s, to, nullptr); }

int launch_rscale_field(mlegs_field *f, int divide);   // ops_field.cu

// The vector operations' entry: scalars in PPP (forward, r*u fused into the loads, inputs read from src[]) or on
// their way to PPP (backward, u/r fused into the stores).  When the batched register-FFT path cannot be taken the
// same thing is done with explicit copies and rscale passes around the plain transforms.
int trans_many_scaled(int n, mlegs_field *const *s, const char *to, unsigned rs_mask, const void *const *src) {
  Context &c = ctx();
  TransOpts o;
  o.rs_mask = rs_mask;
  if (n > MLEGS_MAXB) return fail(MLEGS_E_ARG, "trans_many: too many scalars");
  for (int i = 0; i < n && src; ++i) o.src[i] = src[i];
  const int cur = n > 0 ? space_id(s[0]->space) : -1, dst = space_id(to);
  const bool forward = cur == 0 && dst >= 1, backward = dst == 0 && cur >= 1;
  bool fused = n >= 2 && c.p.np > 1 && fft_reg_supported(c.plan_p.n) && (forward || backward);
  if (fused && c.nranks > 1) fused = (c.nrh % 2 == 0) && !(dst == 0 && cur == 1) && dist_window_batch() >= n;
  for (int i = 1; i < n && fused; ++i)
    fused = space_id(s[i]->space) == cur && s[i]->nrchop_offset == s[0]->nrchop_offset &&
            s[i]->npchop_offset == s[0]->npchop_offset && s[i]->nzchop_offset == s[0]->nzchop_offset;
  if (fused) return trans_many_opt(n, s, to, &o);
  // explicit form
  for (int i = 0; i < n; ++i) {
    if (o.src[i] && o.src[i] != s[i]->e) {
      size_t ne = (size_t)s[i]->loc_sz[0] * s[i]->loc_sz[1] * s[i]->loc_sz[2];
      CUDA_TRY(cudaMemcpyAsync(s[i]->e, o.src[i], ne * sizeof(cplx), cudaMemcpyDeviceToDevice, (cudaStream_t)c.stream));
    }
    if (forward && ((rs_mask >> i) & 1u)) MLEGS_TRY(launch_rscale_field(s[i], 0));
  }
  MLEGS_TRY(trans_many_opt(n, s, to, nullptr));
  for (int i = 0; i < n; ++i)
    if (backward && ((rs_mask >> i) & 1u)) MLEGS_TRY(launch_rscale_field(s[i], 1));
  return MLEGS_OK;
}

static int trans_many_opt(int n, mlegs_field *const *s, const char *to, const TransOpts *opt) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  if (n <= 0) return MLEGS_OK;
  const int dst = space_id(to);
  if (dst < 0) return fail(MLEGS_E_ARG, "trans: only taking PPP, PFP, FFP and FFF for spectral transformation");
  bool uniform = true;
  for (int i = 0; i < n; ++i) {
    if (space_id(s[i]->space) < 0)
      return fail(MLEGS_E_ARG, "trans: scalar space info corrupted (only accepting PPP, PFP, FFP and FFF)");
    uniform = uniform && space_id(s[i]->space) == space_id(s[0]->space) &&
              s[i]->nrchop_offset == s[0]->nrchop_offset && s[i]->npchop_offset == s[0]->npchop_offset &&
              s[i]->nzchop_offset == s[0]->nzchop_offset;
    for (int j = 0; j < i; ++j)
      if (s[i]->e == s[j]->e) return fail(MLEGS_E_ARG, "trans_many: the same scalar appears twice");
  }
  // mixed states: nothing to share
  if (n == 1 || !uniform) {
    if (opt) return fail(MLEGS_E_STATE, "trans_many: fused options need scalars of one state");
    for (int i = 0; i < n; ++i) MLEGS_TRY(trans_impl(s[i], to));
    return MLEGS_OK;
  }
  const int cur = space_id(s[0]->space);
  if (cur == dst) return MLEGS_OK;
  int group = MLEGS_MAXB;
  if (c.nranks > 1 && c.p.np > 1) {
    // several ranks: the fused exchanges carry dist_window_batch() scalars per epoch; paths that need the stand-alone
    // exchange (odd lengths, a backward transform that starts at PFP) go one scalar at a time
    const bool crosses = (cur == 0 && dst >= 1) || (dst == 0 && cur >= 1);
    const bool fused_ok = fft_reg_supported(c.plan_p.n) && (c.nrh % 2 == 0) && !(dst == 0 && cur == 1);
    group = std::min(group, dist_window_batch());
    if (crosses && (!fused_ok || group < 2)) {
      if (opt) return fail(MLEGS_E_STATE, "trans_many: fused options need the fused exchange");
      for (int i = 0; i < n; ++i) MLEGS_TRY(trans_impl(s[i], to));
      return MLEGS_OK;
    }
  }
  {   // precondition checks of rtrans_* once, before anything is launched
    LegArgs a;
    MLEGS_TRY(rtrans_args(s[0], cur < dst ? "rtrans_forward" : "rtrans_backward", &a));
  }
  for (int i0 = 0; i0 < n; i0 += group) MLEGS_TRY(trans_group(std::min(group, n - i0), s + i0, cur, dst, opt, i0));
  return MLEGS_OK;
}

}  // namespace mlegs

using namespace mlegs;

extern "C" {

int mlegs_b200_trans(mlegs_field *s, const char to[3]) { return trans_impl(s, to); }

int mlegs_b200_trans_many(int n, mlegs_field *const *s, const char to[3]) { return trans_many_impl(n, s, to); }

int mlegs_b200_trans_host(void *host_e, const char from[3], const char to[3], double ln) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  cudaStream_t st = (cudaStream_t)c.stream;
  static mlegs_field f;   // staging scalar, allocated on first use
  static size_t f_bytes = 0;
  if (space_id(from) < 0)
    return fail(MLEGS_E_ARG, "trans: scalar space info corrupted (only accepting PPP, PFP, FFP and FFF)");
  if (!f.e || f_bytes != c.field_bytes) {
    if (f.e) cudaFree(f.e);
    CUDA_TRY(cudaMalloc(&f.e, c.field_bytes));
    f_bytes = c.field_bytes;
  }
  field_set_layout(&f, space_id(from) == 0);
  f.nrchop_offset = f.npchop_offset = f.nzchop_offset = 0;
  set_space(&f, space_id(from));
  f.ln = ln;
  size_t n = (size_t)f.loc_sz[0] * f.loc_sz[1] * f.loc_sz[2];
  CUDA_TRY(cudaMemcpyAsync(f.e, host_e, n * sizeof(cplx), cudaMemcpyHostToDevice, st));
  MLEGS_TRY(trans_impl(&f, to));
  n = (size_t)f.loc_sz[0] * f.loc_sz[1] * f.loc_sz[2];
  CUDA_TRY(cudaMemcpyAsync(host_e, f.e, n * sizeof(cplx), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return MLEGS_OK;
}

// Batched host entry: the n host arrays are independent scalars (e.g. the three components a tp2vec caller
// transforms back to back).  The scalars are staged in GROUPS of up to MLEGS_MAXB: a group goes through
// mlegs_b200_trans_many -- one launch per stage for the whole group and, on several ranks, ONE fused exchange and one
// barrier per group instead of one per scalar -- while three sets of staging scalars and two copy streams keep PCIe
// busy in both directions: the H2D copies of group g+1 and the D2H copies of group g-1 overlap the transform of group g.
int mlegs_b200_trans_host_batch(int n, void *const *host_e, const char from[3], const char to[3], const double *ln) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  if (space_id(from) < 0)
    return fail(MLEGS_E_ARG, "trans: scalar space info corrupted (only accepting PPP, PFP, FFP and FFF)");
  if (space_id(to) < 0) return fail(MLEGS_E_ARG, "trans: only taking PPP, PFP, FFP and FFF for spectral transformation");
  cudaStream_t st = (cudaStream_t)c.stream;
  const int NB = 3;
  static mlegs_field f[NB][MLEGS_MAXB];
  static size_t f_bytes = 0;
  static int f_group = 0;
  static cudaStream_t s_in = nullptr, s_out = nullptr;
  static cudaEvent_t in_done[NB], comp_done[NB], out_done[NB];
  if (!s_in) {
    CUDA_TRY(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    for (int b = 0; b < NB; ++b) {
      CUDA_TRY(cudaEventCreateWithFlags(&in_done[b], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&comp_done[b], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&out_done[b], cudaEventDisableTiming));
    }
  }
  // group size: what one exchange epoch carries, within 6 GB of staging memory, at most a third of the batch so that
  // the three-deep pipeline has something to overlap
  // One rank: PCIe is the bottleneck (the transforms of a field take a sixth of its two copies), so the pipeline runs
  // one field at a time and its fill and drain cost one field each (2.15 -> 2.5 GDOF/s at 128^3 against groups of 2).
  // Several ranks: a group shares one exchange + barrier, so it is as large as an epoch carries while leaving three
  // groups to overlap.
  int group = c.nranks > 1 ? std::min({MLEGS_MAXB, dist_window_batch(), (n + NB - 1) / NB}) : 1;
  group = (int)std::max<size_t>(1, std::min<size_t>(std::max(group, 1), ((size_t)6 << 30) / (NB * c.field_bytes)));
  if (f_bytes != c.field_bytes || f_group < group) {
    CUDA_TRY(cudaStreamSynchronize(s_in));
    CUDA_TRY(cudaStreamSynchronize(s_out));
    for (int b = 0; b < NB; ++b)
      for (int q = 0; q < MLEGS_MAXB; ++q) {
        if (f[b][q].e) cudaFree(f[b][q].e);
        f[b][q].e = nullptr;
        if (q < group) CUDA_TRY(cudaMalloc(&f[b][q].e, c.field_bytes));
      }
    f_bytes = c.field_bytes;
    f_group = group;
  }
  // the copy streams start after everything already queued on the compute stream
  CUDA_TRY(cudaEventRecord(comp_done[0], st));
  CUDA_TRY(cudaStreamWaitEvent(s_in, comp_done[0], 0));
  int g = 0;
  for (int j0 = 0; j0 < n; j0 += group, ++g) {
    const int b = g % NB;
    const int cnt = std::min(group, n - j0);
    if (g >= NB) CUDA_TRY(cudaStreamWaitEvent(s_in, out_done[b], 0));   // staging set b is free again
    mlegs_field *ptrs[MLEGS_MAXB];
    for (int q = 0; q < cnt; ++q) {
      mlegs_field &fq = f[b][q];
      field_set_layout(&fq, space_id(from) == 0);
      fq.nrchop_offset = fq.npchop_offset = fq.nzchop_offset = 0;
      set_space(&fq, space_id(from));
      fq.ln = ln ? ln[j0 + q] : 0.0;
      const size_t ne = (size_t)fq.loc_sz[0] * fq.loc_sz[1] * fq.loc_sz[2];
      CUDA_TRY(cudaMemcpyAsync(fq.e, host_e[j0 + q], ne * sizeof(cplx), cudaMemcpyHostToDevice, s_in));
      ptrs[q] = &fq;
    }
    CUDA_TRY(cudaEventRecord(in_done[b], s_in));
    CUDA_TRY(cudaStreamWaitEvent(st, in_done[b], 0));
    MLEGS_TRY(trans_many_impl(cnt, ptrs, to));
    CUDA_TRY(cudaEventRecord(comp_done[b], st));
    CUDA_TRY(cudaStreamWaitEvent(s_out, comp_done[b], 0));
    for (int q = 0; q < cnt; ++q) {
      mlegs_field &fq = f[b][q];
      const size_t ne = (size_t)fq.loc_sz[0] * fq.loc_sz[1] * fq.loc_sz[2];
      CUDA_TRY(cudaMemcpyAsync(host_e[j0 + q], fq.e, ne * sizeof(cplx), cudaMemcpyDeviceToHost, s_out));
    }
    CUDA_TRY(cudaEventRecord(out_done[b], s_out));
  }
  CUDA_TRY(cudaStreamSynchronize(s_out));
  CUDA_TRY(cudaStreamSynchronize(st));
  return MLEGS_OK;
}

}  // extern "C"
