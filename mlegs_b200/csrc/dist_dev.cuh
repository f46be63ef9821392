// Device-side pieces of the peer-memory exchange (see dist.cu): the peer table passed to kernels, the addressing of
// a put, and the system-scope barrier that ends every kernel that wrote into peers' windows.
#pragma once
#include "kernels.h"

namespace mlegs {

#define DIST_MAX_RANKS 16
#define DIST_FLAG_BYTES 4096
#define DIST_RED_DOUBLES 16384                      // per rank and parity (>= 2 nz)
// A rank may legitimately be seconds late to an exchange (a checkpoint write, a first-use factorisation, the host table
// build, Python GC), so waiting never gives up with a wrong answer: the poll backs off and, after ~10 minutes of
// clock64 ticks (a dead peer), traps -- the context dies and every later CUDA call of this rank reports it.
#define DIST_SPIN_LIMIT (1ll << 40)

struct WinHeader {                                  // lives at the start of every window
  unsigned long long arrive[DIST_MAX_RANKS];        // data barrier: epoch published by rank q
  unsigned long long red_arrive[DIST_MAX_RANKS];    // reduction barrier
};

static inline size_t win_red_offset() { return DIST_FLAG_BYTES; }
static inline size_t win_data_offset() {
  return DIST_FLAG_BYTES + (size_t)2 * DIST_MAX_RANKS * DIST_RED_DOUBLES * sizeof(double);
}
static inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

struct DistState {
  bool attached = false;
  void *base[DIST_MAX_RANKS] = {nullptr};           // window base of every rank as mapped here
  unsigned long long epoch = 0, red_epoch = 0;
  unsigned int *d_ctr = nullptr;                    // local CTA counter (last-block detection)
  size_t wstride = 0;                               // bytes of one W buffer (wbatch scalars)
  size_t fbytes = 0;                                // bytes of one scalar's slab inside W
  int wbatch = 1;                                   // scalars one exchange epoch can carry
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.global.release.sys.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.global.acquire.sys.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Destination of local element (i, j, k) of rank `me` in an exchange: owning rank q and linear index inside
// q's new local block.  Shared by the put kernels and the host-side plan (mlegs_b200_dist_put_map), which the
// CPU tests check against the subarray semantics of dist:395-504.  The azimuthal columns are owned cyclically:
// rank q holds m = q, q + P, ... as its local columns 0, 1, ...
// dir 0: (2,1) exchange, src (r_loc, npdim, nz) -> (nrdim, m_cnt[q], nz): column m goes to rank m mod P, local
//        column m / P.
// dir 1: (1,2) exchange, src (nrdim, m_loc, nz) -> (r_cnt[q], npdim, nz).  natural != 0: local column j lands at its
//        global position me + P j (the user-visible s%exchange).  natural == 0: the transit layout of the exchanges
//        fused into a transform -- columns grouped by source rank, rank `me`'s at m_off[me] .. -- which keeps the
//        columns a rank ships to a peer contiguous (long NVLink runs); the azimuthal FFT that follows reads its point
//        m at window column m_off[m mod P] + m / P.
__host__ __device__ inline void slab_put_index(int dir, int me, int nranks, const int *r_cnt, const int *r_off,
                                               const int *m_cnt, const int *m_off, int nrdim, int npdim, int i, int j,
                                               int k, int *q_out, size_t *dst_out, int natural = 0) {
  int q = 0;
  if (dir == 0) {
    q = j % nranks;
    *dst_out = ((size_t)k * m_cnt[q] + j / nranks) * nrdim + r_off[me] + i;
  } else {
    while (q + 1 < nranks && i >= r_off[q + 1]) ++q;
    const int col = natural ? me + nranks * j : m_off[me] + j;
    *dst_out = ((size_t)k * npdim + col) * r_cnt[q] + (i - r_off[q]);
  }
  *q_out = q;
}

// Staged (1,2) exchange: element (i, j, k) of the local (nrdim, m_cnt[me], nz) block inside a LOCAL staging buffer laid
// out in destination order -- one region per owning rank q, each region exactly the piece of q's new block this rank
// contributes ((r_cnt[q], m_cnt[me]) per plane), so that shipping it is nz contiguous runs per peer.
__host__ __device__ inline size_t slab_stage_index(int me, int nranks, const int *r_cnt, const int *r_off, const int *m_cnt,
                                                   int nz, int i, int j, int k) {
  int q = 0;
  while (q + 1 < nranks && i >= r_off[q + 1]) ++q;
  return (size_t)r_off[q] * m_cnt[me] * nz + ((size_t)k * m_cnt[me] + j) * r_cnt[q] + (i - r_off[q]);
}

struct PeerTable {
  void *base[DIST_MAX_RANKS];
  int r_cnt[DIST_MAX_RANKS], r_off[DIST_MAX_RANKS], m_cnt[DIST_MAX_RANKS], m_off[DIST_MAX_RANKS];
  int rank, nranks;
  size_t data_off;      // byte offset of W[parity] inside a window
  size_t fstride;       // bytes between the scalars of a batched exchange inside W
  unsigned long long epoch;
  unsigned int *ctr;
  int *flag;            // device error flags ([2] = exchange timeout)
};

// publish `epoch` to all peers and wait for theirs; called by the first warp of the last CTA (one lane per peer, so
// the remote stores and the polls of all peers are in flight together instead of one NVLink round trip each)
__device__ __forceinline__ void barrier_publish_wait(const PeerTable &t, bool reduction) {
  const int q = threadIdx.x;
  if (q < t.nranks) {
    WinHeader *h = reinterpret_cast<WinHeader *>(t.base[q]);
    st_release_sys(reduction ? &h->red_arrive[t.rank] : &h->arrive[t.rank], t.epoch);
    WinHeader *me = reinterpret_cast<WinHeader *>(t.base[t.rank]);
    const unsigned long long *p = reduction ? &me->red_arrive[q] : &me->arrive[q];
    const long long t0 = clock64();
    unsigned backoff = 0;
    while (ld_acquire_sys(p) < t.epoch) {
      if (backoff < 4096u) {
        ++backoff;                       // the first microseconds: poll at full rate (the common case)
      } else {
        __nanosleep(1000);               // a late peer: one poll per microsecond keeps NVLink and the LSU free
        if (clock64() - t0 > DIST_SPIN_LIMIT) {
          atomicOr(t.flag + 2, 1);
          __trap();                      // never continue past an incomplete barrier
        }
      }
    }
  }
  __syncwarp();
}

// Ends a kernel that wrote into peers' windows: make this CTA's puts visible system-wide, count the CTA, and let the
// last CTA publish the epoch to every peer and wait for theirs.  Call with all threads of the CTA.
__device__ __forceinline__ void dist_finish_put(const PeerTable &t, unsigned int nblocks) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < 32) {           // first warp (every kernel that calls this has >= 32 threads)
    unsigned int prev = 0;
    if (threadIdx.x == 0) prev = atomicAdd(t.ctr, 1u);
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev == nblocks - 1) {
      if (threadIdx.x == 0) {
        *t.ctr = 0;
        __threadfence_system();
      }
      __syncwarp();
      barrier_publish_wait(t, false);
    }
  }
}

}  // namespace mlegs
