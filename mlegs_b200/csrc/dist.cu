// scalar_exchange (/root/reference/src/submodules/mlegs_scalar_dist.f90:6-67) in the slab layout.
#include "kernels.h"

namespace mlegs {

// Move the local block from `src` to `dst` so that axis_new becomes local and axis_old distributed.
int exchange_slab(mlegs_field *s, int axis_old, int axis_new, const void *src, void *dst) {
  Context &c = ctx();
  if (s->axis_comm[axis_old - 1] != 0)
    return fail(MLEGS_E_COMM,
                "ERROR: scalar_exchange requires the data to be non-distributed along the old dimension");
  if (c.nranks == 1) {
    if (src != dst) {
      size_t n = (size_t)s->loc_sz[0] * s->loc_sz[1] * s->loc_sz[2];
      CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(cplx), cudaMemcpyDeviceToDevice, (cudaStream_t)c.stream));
    }
    s->axis_comm[axis_old - 1] = s->axis_comm[axis_new - 1];
    s->axis_comm[axis_new - 1] = 0;
    return MLEGS_OK;
  }
  return fail(MLEGS_E_COMM, "scalar_exchange: multi-rank exchange window is not attached");
}

}  // namespace mlegs

using namespace mlegs;

extern "C" {

int mlegs_b200_exchange(mlegs_field *s, int axis_old, int axis_new) {
  Context &c = ctx();
  if (!c.ready) return fail(MLEGS_E_STATE, "mlegs_b200: transformation kit is not initialized");
  if (axis_old < 1 || axis_old > 3 || axis_new < 1 || axis_new > 3 || axis_old == axis_new)
    return fail(MLEGS_E_COMM, "scalar_exchange: invalid axes");
  if (s->axis_comm[axis_new - 1] == 0) {
    // nothing is distributed along axis_new: the reference would index comm_grps(0); treat as a no-op
    return MLEGS_OK;
  }
  return exchange_slab(s, axis_old, axis_new, s->e, s->e);
}

int mlegs_b200_dist_window(void **dev_ptr, size_t *bytes, unsigned char handle64[64]) {
  (void)dev_ptr; (void)bytes; (void)handle64;
  return fail(MLEGS_E_STATE, "mlegs_b200_dist_window: not implemented yet");
}
int mlegs_b200_dist_attach(const unsigned char *h) {
  (void)h;
  return fail(MLEGS_E_STATE, "mlegs_b200_dist_attach: not implemented yet");
}
int mlegs_b200_dist_detach(void) { return MLEGS_OK; }

}  // extern "C"
