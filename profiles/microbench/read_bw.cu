// Read-only streaming bandwidth on this B200 at the sizes the sweep kernels see
// (calibrates what "HBM roofline" means for a 25-100 MB read-mostly pass).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_bw read_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

template <int U>
__global__ void __launch_bounds__(256) read_kernel(const uint4* __restrict__ p, size_t n_vec, uint32_t* out) {
  size_t i = (size_t)blockIdx.x * 256 * U + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * 256 * U;
  uint32_t acc = 0;
  for (; i < n_vec; i += stride) {
    uint4 v[U];
#pragma unroll
    for (int j = 0; j < U; j++) v[j] = (i + j * 256 < n_vec) ? ldg_stream(p + i + j * 256) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < U; j++) acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
  }
  if (acc == 0x12345678u) out[0] = acc;  // never true in practice; keeps the loads alive
}

int main() {
  const size_t sizes_mb[] = {8, 25, 50, 100, 400, 1600};
  const int copies_target_mb = 400;  // rotate so that every launch reads cold lines
  uint32_t* out;
  cudaMalloc(&out, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int sms = 148;
  printf("size_MB,grid,unroll,us_per_launch,GB_s\n");
  for (size_t mb : sizes_mb) {
    size_t bytes = mb << 20;
    int copies = (int)((copies_target_mb + mb - 1) / mb);
    if (copies < 1) copies = 1;
    if (copies > 16) copies = 16;
    std::vector<uint4*> bufs(copies);
    for (auto& b : bufs) {
      cudaMalloc(&b, bytes);
      cudaMemset(b, 1, bytes);
    }
    size_t n_vec = bytes / 16;
    for (int mode = 0; mode < 3; mode++) {
      const int U = 4;
      size_t want = (n_vec + 256 * U - 1) / (256 * U);
      unsigned grid = mode == 0 ? (unsigned)want : mode == 1 ? sms * 8 : sms * 4;
      if (grid > want) grid = (unsigned)want;
      const int iters = 50;
      for (int w = 0; w < 5; w++) read_kernel<U><<<grid, 256>>>(bufs[w % copies], n_vec, out);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      for (int it = 0; it < iters; it++) read_kernel<U><<<grid, 256>>>(bufs[it % copies], n_vec, out);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double us = ms * 1e3 / iters;
      printf("%zu,%u,%d,%.2f,%.0f\n", mb, grid, U, us, bytes / us / 1e3);
    }
    for (auto& b : bufs) cudaFree(b);
  }
  return 0;
}
