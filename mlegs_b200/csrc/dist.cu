// scalar_exchange (/root/reference/src/submodules/mlegs_scalar_dist.f90:6-67, 468-504) in the slab layout,
// as one-sided puts over NVLink peer memory instead of MPI_Alltoallw with derived datatypes.
//
// Every rank owns one CUDA-IPC-exported allocation ("window"):
//     [ flags | reduction slots (2 parities) | W0 | W1 ]
// W0/W1 are field-sized receive buffers used alternately (epoch parity).  An exchange is ONE kernel:
// each rank scatters its local block straight into the peers' W[epoch & 1] with coalesced 16-byte stores
// (the transposition the MPI datatypes describe is done by the addressing), fences at system scope, and the
// last CTA publishes `epoch` into every peer's arrive[] slot and waits until all peers have published
// theirs.  Double buffering makes that single barrier sufficient: a peer can only write W[p] again two
// epochs later, after it has seen this rank's signal of the epoch in between, which this rank sends (in
// stream order) after its readers of W[p] have finished.
//
// The tiny all-reduces of the reference (SVV energies ops:117-118, calcat ops:265,302, ln ops:394,657,748)
// use the same mechanism on the reduction slots and sum in rank order on every rank, so all ranks hold
// bit-identical results.
#include <algorithm>
#include <cstring>
#include <string>

#include "dist_dev.cuh"

namespace mlegs {

static DistState g_dist;

// dir 0: (2,1) exchange, src (r_loc, npdim, nz) -> peers' (nrdim, m_cnt[q], nz)   [phi local -> r local]
// dir 1: (1,2) exchange, src (nrdim, m_loc, nz) -> peers' (r_cnt[q], npdim, nz)   [r local -> phi local]
__global__ void __launch_bounds__(256) exchange_put_kernel(PeerTable t, const cplx *__restrict__ src, int dir, int nrdim,
                                                           int npdim, int nz) {
  const int me = t.rank;
  const int rows = dir == 0 ? t.r_cnt[me] : nrdim;
  const int cols = dir == 0 ? npdim : t.m_cnt[me];
  const size_t n = (size_t)rows * cols * nz;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % rows);
    const size_t r = idx / rows;
    const int j = (int)(r % cols);
    const int k = (int)(r / cols);
    int q;
    size_t dst;
    slab_put_index(dir, me, t.nranks, t.r_cnt, t.r_off, t.m_cnt, t.m_off, nrdim, npdim, i, j, k, &q, &dst, 1);
    cplx *w = reinterpret_cast<cplx *>(reinterpret_cast<char *>(t.base[q]) + t.data_off);
    w[dst] = src[idx];
  }
  dist_finish_put(t, gridDim.x);
}

// Second half of the staged (1,2) exchange: the Legendre synthesis left every scalar's rows in a local staging buffer in
// destination order (slab_stage_index); this kernel ships them -- per (scalar, peer, plane) one contiguous run of
// m_cnt[me] * r_cnt[q] elements, 16 bytes per lane, whole warps on consecutive addresses -- and ends with the exchange
// barrier.  Rank `me` starts with peer me+1, so at any moment the ranks target different peers.
struct ShipSrc {
  const cplx *p[MLEGS_MAXB];
};
__global__ void __launch_bounds__(256) slab_ship_kernel(PeerTable t, ShipSrc src, int nfld, int nz, int npdim) {
  const int me = t.rank, P = t.nranks, mc = t.m_cnt[me];
  const long long nruns = (long long)nfld * P * nz;
  for (long long run = blockIdx.x; run < nruns; run += gridDim.x) {
    const int k = (int)(run % nz);
    const long long rest = run / nz;
    const int q = (int)((me + 1 + rest % P) % P);
    const int fld = (int)(rest / P);
    const int len = mc * t.r_cnt[q];
    const cplx *s = src.p[fld] + (size_t)t.r_off[q] * mc * nz + (size_t)k * len;
    cplx *d = reinterpret_cast<cplx *>(reinterpret_cast<char *>(t.base[q]) + t.data_off + fld * t.fstride) +
              ((size_t)k * npdim + t.m_off[me]) * t.r_cnt[q];
    for (int e = threadIdx.x; e < len; e += blockDim.x) d[e] = s[e];
  }
  dist_finish_put(t, gridDim.x);
}

// The same runs moved by the TMA unit: one thread per CTA streams chunks of a run through a ring of shared-memory
// stages -- cp.async.bulk global -> shared (mbarrier completion), cp.async.bulk shared -> the peer's window (bulk-group
// completion) -- so the NVLink writes leave as whole bursts instead of one 16-byte store per thread, and the SM's LSU
// and registers stay out of the data path.
#define SHIP_STAGES 4
#define SHIP_CHUNK 8192
__global__ void __launch_bounds__(32) slab_ship_tma_kernel(PeerTable t, ShipSrc src, int nfld, int nz, int npdim) {
  __shared__ __align__(128) unsigned char buf[SHIP_STAGES][SHIP_CHUNK];
  __shared__ unsigned long long full[SHIP_STAGES];
  const int me = t.rank, P = t.nranks, mc = t.m_cnt[me];
  const long long nruns = (long long)nfld * P * nz;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SHIP_STAGES; ++s) {
      unsigned a = (unsigned)__cvta_generic_to_shared(&full[s]);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(a) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    // chunk c of this CTA: runs blockIdx.x, blockIdx.x + gridDim.x, ..., each cut into pieces of SHIP_CHUNK bytes
    long long run_l = blockIdx.x, run_s = blockIdx.x;     // run of the next chunk to load / to store
    unsigned off_l = 0, off_s = 0;                        // byte offset inside that run
    auto locate = [&](long long run, const char **sp, char **dp, unsigned *bytes) {
      const int k = (int)(run % nz);
      const long long rest = run / nz;
      const int q = (int)((me + 1 + rest % P) % P);
      const int fld = (int)(rest / P);
      const int len = mc * t.r_cnt[q];
      *sp = reinterpret_cast<const char *>(src.p[fld] + (size_t)t.r_off[q] * mc * nz + (size_t)k * len);
      *dp = reinterpret_cast<char *>(reinterpret_cast<cplx *>(reinterpret_cast<char *>(t.base[q]) + t.data_off +
                                                               fld * t.fstride) +
                                     ((size_t)k * npdim + t.m_off[me]) * t.r_cnt[q]);
      *bytes = (unsigned)len * 16u;
    };
    auto load = [&](int i) {            // chunk i -> stage i % SHIP_STAGES; false when the CTA has no more chunks
      while (run_l < nruns) {
        const char *sp;
        char *dp;
        unsigned bytes;
        locate(run_l, &sp, &dp, &bytes);
        if (off_l >= bytes) {
          run_l += gridDim.x;
          off_l = 0;
          continue;
        }
        const unsigned n = min(bytes - off_l, (unsigned)SHIP_CHUNK);
        const int s = i % SHIP_STAGES;
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&full[s]);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&buf[s][0]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(n) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                     "l"(sp + off_l), "r"(n), "r"(bar)
                     : "memory");
        off_l += n;
        return true;
      }
      return false;
    };
    int loaded = 0;
    for (; loaded < SHIP_STAGES - 1; ++loaded)
      if (!load(loaded)) break;
    bool more = loaded == SHIP_STAGES - 1;
    for (int i = 0; i < loaded; ++i) {
      // store chunk i
      const char *sp;
      char *dp;
      unsigned bytes;
      for (;;) {
        locate(run_s, &sp, &dp, &bytes);
        if (off_s < bytes) break;
        run_s += gridDim.x;
        off_s = 0;
      }
      const unsigned n = min(bytes - off_s, (unsigned)SHIP_CHUNK);
      const int s = i % SHIP_STAGES;
      const unsigned bar = (unsigned)__cvta_generic_to_shared(&full[s]);
      const unsigned parity = (unsigned)((i / SHIP_STAGES) & 1);
      asm volatile(
          "{\n.reg .pred p;\nSHIP_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra SHIP_DONE;\nbra "
          "SHIP_WAIT;\nSHIP_DONE:\n}\n" ::"r"(bar),
          "r"(parity)
          : "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dp + off_s),
                   "r"((unsigned)__cvta_generic_to_shared(&buf[s][0])), "r"(n)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      off_s += n;
      // refill the stage whose store was committed one iteration ago (chunk i - 1's): at most this iteration's
      // group may still be reading shared memory
      if (more) {
        asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
        if (load(loaded)) ++loaded; else more = false;
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
  }
  dist_finish_put(t, gridDim.x);
}

int launch_slab_ship(const PeerTable &t, const FieldBatch &fb, cudaStream_t st) {
  Context &c = ctx();
  ShipSrc src;
  for (int i = 0; i < fb.n; ++i) src.p[i] = fb.out[i];
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long nruns = (long long)fb.n * c.nranks * c.nzdim;
  static const char *mode = getenv("MLEGS_SHIP");        // "simt" / "tma": A/B timing
  const bool tma = mode ? mode[0] == 't' : false;
  prof_begin("exchange_12_ship", st);
  if (tma) {
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(nruns, (long long)sms * 4));
    slab_ship_tma_kernel<<<grid, 32, 0, st>>>(t, src, fb.n, c.nzdim, c.npdim);
  } else {
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(nruns, (long long)sms * 8));
    slab_ship_kernel<<<grid, 256, 0, st>>>(t, src, fb.n, c.nzdim, c.npdim);
  }
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// sum over ranks of `n` doubles, in rank order, result on every rank (one CTA)
__global__ void allreduce_small_kernel(PeerTable t, double *inout, int n, size_t red_off) {
  for (int q = 0; q < t.nranks; ++q) {
    double *slot = reinterpret_cast<double *>(reinterpret_cast<char *>(t.base[q]) + red_off) +
                   (size_t)t.rank * DIST_RED_DOUBLES;
    for (int i = threadIdx.x; i < n; i += blockDim.x) slot[i] = inout[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < 32) barrier_publish_wait(t, true);
  __syncthreads();
  const double *mine = reinterpret_cast<const double *>(reinterpret_cast<char *>(t.base[t.rank]) + red_off);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int q = 0; q < t.nranks; ++q) acc += *((volatile const double *)&mine[(size_t)q * DIST_RED_DOUBLES + i]);
    inout[i] = acc;
  }
}

static void fill_table(PeerTable *t) {
  Context &c = ctx();
  memset(t, 0, sizeof(*t));
  for (int q = 0; q < c.nranks; ++q) {
    t->base[q] = g_dist.base[q];
    t->r_cnt[q] = c.r_cnt[q];
    t->r_off[q] = c.r_off[q];
    t->m_cnt[q] = c.m_cnt[q];
    t->m_off[q] = c.m_off[q];
  }
  t->rank = c.rank;
  t->nranks = c.nranks;
  t->ctr = g_dist.d_ctr;
  t->flag = c.d_flag;
  t->fstride = g_dist.fbytes;
}

// scalars one exchange epoch can carry (mlegs_b200_trans_many on several ranks)
int dist_window_batch() { return g_dist.attached ? g_dist.wbatch : 1; }
size_t dist_field_stride() { return g_dist.fbytes; }

bool dist_active() { return ctx().nranks > 1 && g_dist.attached; }

// Start an exchange epoch whose puts are issued by some other kernel (the fused FFT / Legendre stores): fills the
// peer table (bases already point at W[epoch & 1]) and returns where this rank's new block will land.
int dist_begin_put(PeerTable *t, void **landed) {
  Context &c = ctx();
  if (!g_dist.attached) return fail(MLEGS_E_COMM, "scalar_exchange: multi-rank exchange window is not attached");
  fill_table(t);
  t->epoch = ++g_dist.epoch;
  t->data_off = win_data_offset() + (size_t)(t->epoch & 1) * g_dist.wstride;
  *landed = reinterpret_cast<char *>(g_dist.base[c.rank]) + t->data_off;
  return MLEGS_OK;
}

// Sum `n` device doubles over all ranks (no-op on one rank).
int dist_allreduce(double *d_inout, int n) {
  Context &c = ctx();
  if (c.nranks == 1) return MLEGS_OK;
  if (!g_dist.attached) return fail(MLEGS_E_COMM, "mlegs_b200: multi-rank run without attached exchange windows");
  if (n > DIST_RED_DOUBLES) return fail(MLEGS_E_ARG, "dist_allreduce: too many values");
  PeerTable t;
  fill_table(&t);
  t.epoch = ++g_dist.red_epoch;
  size_t red_off = win_red_offset() + (size_t)(t.epoch & 1) * DIST_MAX_RANKS * DIST_RED_DOUBLES * sizeof(double);
  cudaStream_t st = (cudaStream_t)c.stream;
  prof_begin("allreduce_small", st);
  allreduce_small_kernel<<<1, 256, 0, st>>>(t, d_inout, n, red_off);
  prof_end(st);
  KERNEL_CHECK();
  return MLEGS_OK;
}

// Move the local block at `src` so that axis_new becomes local and axis_old distributed.  On return *landed
// points at the buffer holding the new local block (the receive window on several ranks, `src` on one).
int exchange_slab(mlegs_field *s, int axis_old, int axis_new, const void *src, void **landed) {
  Context &c = ctx();
  *landed = const_cast<void *>(src);
  if (c.nranks == 1) return MLEGS_OK;
  if (!g_dist.attached) return fail(MLEGS_E_COMM, "scalar_exchange: multi-rank exchange window is not attached");
  int dir;
  if (axis_old == 2 && axis_new == 1)
    dir = 0;
  else if (axis_old == 1 && axis_new == 2)
    dir = 1;
  else
    return fail(MLEGS_E_COMM, "scalar_exchange: the slab layout only exchanges axes (2,1) and (1,2)");
  PeerTable t;
  fill_table(&t);
  t.epoch = ++g_dist.epoch;
  t.data_off = win_data_offset() + (size_t)(t.epoch & 1) * g_dist.wstride;
  const size_t n = (size_t)s->loc_sz[0] * s->loc_sz[1] * s->loc_sz[2];
  unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
  if (grid == 0) grid = 1;
  cudaStream_t st = (cudaStream_t)c.stream;
  prof_begin(dir == 0 ? "exchange_21" : "exchange_12", st);
  exchange_put_kernel<<<grid, 256, 0, st>>>(t, (const cplx *)src, dir, c.nrdim, c.npdim, c.nzdim);
  prof_end(st);
  KERNEL_CHECK();
  *landed = reinterpret_cast<char *>(g_dist.base[c.rank]) + t.data_off;
  int keep[3] = {s->axis_comm[0], s->axis_comm[1], s->axis_comm[2]};
  field_set_layout(s, dir == 1);          // new local sizes / offsets
  s->axis_comm[0] = keep[0];              // the labels are the caller's business (reference: dist:52-58)
  s->axis_comm[1] = keep[1];
  s->axis_comm[2] = keep[2];
  return MLEGS_OK;
}

// Host-side completion check of everything queued so far, including exchange barriers: a barrier that waited ~10 minutes
// for a dead peer traps (dist_dev.cuh), which surfaces here -- and at every later CUDA call -- as a launch failure.
int dist_check_timeout() {
  Context &c = ctx();
  if (c.nranks == 1) return MLEGS_OK;
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)c.stream);
  if (e != cudaSuccess)
    return fail(MLEGS_E_COMM, std::string("scalar_exchange: exchange barrier failed (a peer rank never arrived): ") +
                                  cudaGetErrorString(e));
  return MLEGS_OK;
}

}  // namespace mlegs

using namespace mlegs;

extern "C" {

int mlegs_b200_exchange(mlegs_field *s, int axis_old, int axis_new) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  if (axis_old < 1 || axis_old > 3 || axis_new < 1 || axis_new > 3 || axis_old == axis_new)
    return fail(MLEGS_E_COMM, "scalar_exchange: invalid axes");
  const int g_new = s->axis_comm[axis_new - 1];
  if (g_new == 0) return MLEGS_OK;   // nothing is distributed along axis_new (the reference would index comm_grps(0))
  if (s->axis_comm[axis_old - 1] != 0)   // dist:16-20
    return fail(MLEGS_E_COMM,
                "ERROR: scalar_exchange requires the data to be non-distributed along the old dimension");
  if (c.nranks > 1 && g_new == 1) {
    // comm_grps(1) holds all the ranks of the slab decomposition: a real all-to-all
    void *landed = nullptr;
    MLEGS_TRY(exchange_slab(s, axis_old, axis_new, s->e, &landed));
    if (landed != s->e) {
      size_t n = (size_t)s->loc_sz[0] * s->loc_sz[1] * s->loc_sz[2];
      CUDA_TRY(cudaMemcpyAsync(s->e, landed, n * sizeof(cplx), cudaMemcpyDeviceToDevice, (cudaStream_t)c.stream));
    }
  }
  // comm_grps(2) (and everything on one rank) is a single-rank group: the exchange is a re-labelling (dist:52-58)
  s->axis_comm[axis_old - 1] = g_new;
  s->axis_comm[axis_new - 1] = 0;
  return MLEGS_OK;
}

int mlegs_b200_dist_m_stride(const mlegs_field *s, int *stride) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  *stride = field_mstride(s);
  return MLEGS_OK;
}

int mlegs_b200_dist_window(void **dev_ptr, size_t *bytes, unsigned char handle64[64]) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  if (c.nranks > DIST_MAX_RANKS) return fail(MLEGS_E_COMM, "mlegs_b200: at most 16 ranks per node are supported");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  static_assert(sizeof(WinHeader) <= DIST_FLAG_BYTES, "header size");
  if (!c.d_window) {
    // W0/W1 carry up to 8 scalars per epoch (batched transforms; one-GPU launches take up to MLEGS_MAXB = 32), within 8 GB per rank
    g_dist.fbytes = align256(c.field_bytes);
    g_dist.wbatch = (int)std::max<size_t>(1, std::min<size_t>(std::min(MLEGS_MAXB, 8), ((size_t)8 << 30) / (2 * g_dist.fbytes)));
    g_dist.wstride = g_dist.fbytes * g_dist.wbatch;
    size_t total = win_data_offset() + 2 * g_dist.wstride;
    CUDA_TRY(cudaMalloc(&c.d_window, total));
    CUDA_TRY(cudaMemset(c.d_window, 0, win_data_offset()));
    CUDA_TRY(cudaMalloc((void **)&g_dist.d_ctr, sizeof(unsigned int)));
    CUDA_TRY(cudaMemset(g_dist.d_ctr, 0, sizeof(unsigned int)));
    CUDA_TRY(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, c.d_window));
  memcpy(handle64, &h, 64);
  if (dev_ptr) *dev_ptr = c.d_window;
  if (bytes) *bytes = win_data_offset() + 2 * g_dist.wstride;
  return MLEGS_OK;
}

int mlegs_b200_dist_attach(const unsigned char *handles64_all_ranks) {
  Context &c = ctx();
  if (!c.ready || !c.d_window) return fail(MLEGS_E_STATE, "mlegs_b200_dist_attach: call mlegs_b200_dist_window first");
  for (int q = 0; q < c.nranks; ++q) {
    if (q == c.rank) {
      g_dist.base[q] = c.d_window;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles64_all_ranks + (size_t)64 * q, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(&g_dist.base[q], h, cudaIpcMemLazyEnablePeerAccess));
  }
  g_dist.epoch = g_dist.red_epoch = 0;
  g_dist.attached = true;
  return MLEGS_OK;
}

int mlegs_b200_dist_detach(void) {
  Context &c = ctx();
  if (g_dist.attached) {
    cudaDeviceSynchronize();
    for (int q = 0; q < c.nranks; ++q)
      if (q != c.rank && g_dist.base[q]) cudaIpcCloseMemHandle(g_dist.base[q]);
  }
  if (c.d_window) cudaFree(c.d_window);
  c.d_window = nullptr;
  if (g_dist.d_ctr) cudaFree(g_dist.d_ctr);
  g_dist = DistState();
  return MLEGS_OK;
}

/* Host-only (no CUDA): the exchange plan of rank `rank` out of `nranks` for global sizes (nrdim, npdim, nz):
 * for every element of the local block, in memory order, the destination rank and the linear index inside
 * that rank's new local block.  Same code as the device put kernel. */
int mlegs_b200_dist_put_map(int dir, int rank, int nranks, int nrdim, int npdim, int nz, int *dst_rank,
                            long long *dst_index) {
  if (nranks < 1 || nranks > DIST_MAX_RANKS || rank < 0 || rank >= nranks || dir < 0 || dir > 2)
    return fail(MLEGS_E_COMM, "mlegs_b200_dist_put_map: bad arguments");
  const int natural = dir == 2 ? 0 : 1;   // dir 2: exchange(1,2) into the transit layout of the fused exchanges
  if (dir == 2) dir = 1;
  int r_cnt[DIST_MAX_RANKS], r_off[DIST_MAX_RANKS], m_cnt[DIST_MAX_RANKS], m_off[DIST_MAX_RANKS];
  for (int q = 0; q < nranks; ++q) {
    decompose(nrdim, nranks, q, &r_cnt[q], &r_off[q]);
    m_cnt[q] = q < npdim ? (npdim - q + nranks - 1) / nranks : 0;   // cyclic: m = q, q + P, ...
    m_off[q] = q == 0 ? 0 : m_off[q - 1] + m_cnt[q - 1];
  }
  const int rows = dir == 0 ? r_cnt[rank] : nrdim;
  const int cols = dir == 0 ? npdim : m_cnt[rank];
  size_t idx = 0;
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < cols; ++j)
      for (int i = 0; i < rows; ++i, ++idx) {
        int q;
        size_t dst;
        slab_put_index(dir, rank, nranks, r_cnt, r_off, m_cnt, m_off, nrdim, npdim, i, j, k, &q, &dst, natural);
        dst_rank[idx] = q;
        dst_index[idx] = (long long)dst;
      }
  return MLEGS_OK;
}

/* Host-only: the staged exchange(1,2) of rank `rank` as two maps, produced by the very addressing code the device runs:
 * stage_index = slab_stage_index (Legendre synthesis epilogue, MODE 2), ship_* = the run arithmetic of slab_ship_kernel. */
int mlegs_b200_dist_stage_map(int rank, int nranks, int nrdim, int npdim, int nz, long long *stage_index,
                              int *ship_rank, long long *ship_index) {
  if (nranks < 1 || nranks > DIST_MAX_RANKS || rank < 0 || rank >= nranks)
    return fail(MLEGS_E_COMM, "mlegs_b200_dist_stage_map: bad arguments");
  int r_cnt[DIST_MAX_RANKS], r_off[DIST_MAX_RANKS], m_cnt[DIST_MAX_RANKS], m_off[DIST_MAX_RANKS];
  for (int q = 0; q < nranks; ++q) {
    decompose(nrdim, nranks, q, &r_cnt[q], &r_off[q]);
    m_cnt[q] = q < npdim ? (npdim - q + nranks - 1) / nranks : 0;   // cyclic: m = q, q + P, ...
    m_off[q] = q == 0 ? 0 : m_off[q - 1] + m_cnt[q - 1];
  }
  const int mc = m_cnt[rank];
  size_t idx = 0;
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < mc; ++j)
      for (int i = 0; i < nrdim; ++i, ++idx)
        stage_index[idx] = (long long)slab_stage_index(rank, nranks, r_cnt, r_off, m_cnt, nz, i, j, k);
  // the ship kernel: per (peer q, plane k) one run of mc * r_cnt[q] elements
  for (int q = 0; q < nranks; ++q)
    for (int k = 0; k < nz; ++k) {
      const int len = mc * r_cnt[q];
      const size_t s0 = (size_t)r_off[q] * mc * nz + (size_t)k * len;
      const size_t d0 = ((size_t)k * npdim + m_off[rank]) * r_cnt[q];
      for (int e = 0; e < len; ++e) {
        ship_rank[s0 + e] = q;
        ship_index[s0 + e] = (long long)(d0 + e);
      }
    }
  return MLEGS_OK;
}

/* sum `n` HOST doubles over all ranks (the app-level MPI_Allreduce of e.g. check_stability,
 * apps/vortical_flow_3d.f90:404, expressed on the library's own peer windows) */
int mlegs_b200_dist_allreduce(double *host_inout, int n) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  if (c.nranks == 1) return MLEGS_OK;
  if (n > Context::RED_DOUBLES) return fail(MLEGS_E_ARG, "mlegs_b200_dist_allreduce: too many values");
  cudaStream_t st = (cudaStream_t)c.stream;
  CUDA_TRY(cudaMemcpyAsync(c.d_red, host_inout, n * sizeof(double), cudaMemcpyHostToDevice, st));
  MLEGS_TRY(dist_allreduce(c.d_red, n));
  CUDA_TRY(cudaMemcpyAsync(host_inout, c.d_red, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return dist_check_timeout();
}

}  // extern "C"
