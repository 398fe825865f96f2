// Entry points whose kernels are not written yet report cudaErrorNotSupported
// (→ LWSE_ERR_CUDA); each is removed from this file when its kernel lands.
#include "lwse_device.cuh"

namespace lwse {
#ifndef LWSE_HAVE_PLACE
int launch_place(const lwse_node_rec*, uint32_t, uint32_t, const lwse_place_req*, uint32_t,
                 const uint32_t*, uint32_t, lwse_place_out*, void*, size_t, uint32_t*, int,
                 cudaStream_t, int* cuda_err) {
  *cuda_err = (int)cudaErrorNotSupported;
  return -1;
}
size_t place_scratch_bytes(uint32_t, uint32_t, uint32_t, uint32_t) { return 256; }
#endif
#ifndef LWSE_HAVE_DS
int launch_ds_sweep(const lwse_ds_tables*, int, cudaStream_t, int* cuda_err) {
  *cuda_err = (int)cudaErrorNotSupported;
  return -1;
}
#endif
#ifndef LWSE_HAVE_SHA1
int launch_sha1(const uint8_t*, const uint32_t*, uint32_t, uint8_t*, int, cudaStream_t, int* cuda_err) {
  *cuda_err = (int)cudaErrorNotSupported;
  return -1;
}
#endif
}  // namespace lwse
