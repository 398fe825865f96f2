// Host-side cost of the CUDA runtime calls a tick makes (B200 box): how long does the CALL take,
// and when does the GPU start the work?  nvcc -O2 -gencode arch=compute_100a,code=sm_100a call_cost.cu -o call_cost
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstring>

__global__ void tiny(unsigned* p) { if (p && threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1; }
__global__ void spin_flag(volatile unsigned* flag) { flag[0] = 1; }

static double now_us() {
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main() {
  cudaStream_t s, s2;
  cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  void *h_plain, *h_mapped, *d;
  const size_t cap = 8u << 20;
  cudaHostAlloc(&h_plain, cap, cudaHostAllocDefault);
  cudaHostAlloc(&h_mapped, cap, cudaHostAllocMapped);
  cudaMalloc(&d, cap);
  memset(h_plain, 1, cap);
  memset(h_mapped, 1, cap);
  unsigned* d_word;
  cudaMalloc(&d_word, 64);
  cudaMemset(d_word, 0, 64);
  volatile unsigned* h_flag;
  cudaHostAlloc((void**)&h_flag, 64, cudaHostAllocMapped);
  unsigned* d_flag;
  cudaHostGetDevicePointer((void**)&d_flag, (void*)h_flag, 0);
  for (int i = 0; i < 20; i++) tiny<<<1, 32, 0, s>>>(d_word);
  cudaDeviceSynchronize();
  const size_t sizes[] = {4096, 65536, 262144, 425984, 1u << 20, 4u << 20};
  for (int mapped = 0; mapped < 2; mapped++) {
    void* h = mapped ? h_mapped : h_plain;
    for (size_t sz : sizes) {
      double call = 0, total = 0, kern_after = 0;
      const int reps = 50;
      for (int r = 0; r < reps; r++) {
        cudaDeviceSynchronize();
        h_flag[0] = 0;
        const double t0 = now_us();
        cudaMemcpyAsync(d, h, sz, cudaMemcpyHostToDevice, s);
        const double t1 = now_us();
        spin_flag<<<1, 1, 0, s2>>>(d_flag);  // another stream: when does a kernel enqueued right after start?
        const double t2 = now_us();
        while (h_flag[0] == 0) {}
        const double t3 = now_us();
        cudaStreamSynchronize(s);
        const double t4 = now_us();
        call += t1 - t0;
        kern_after += t3 - t1;
        total += t4 - t0;
        (void)t2;
      }
      printf("memcpyAsync %s %8zu B: call %.2f us, kernel on another stream visible after +%.2f us, copy done after %.2f us\n",
             mapped ? "mapped " : "default", sz, call / reps, kern_after / reps, total / reps);
    }
  }
  {  // plain kernel launch cost and launch-to-visible latency
    double call = 0, vis = 0;
    const int reps = 200;
    for (int r = 0; r < reps; r++) {
      cudaDeviceSynchronize();
      h_flag[0] = 0;
      const double t0 = now_us();
      spin_flag<<<1, 1, 0, s>>>(d_flag);
      const double t1 = now_us();
      while (h_flag[0] == 0) {}
      const double t2 = now_us();
      call += t1 - t0;
      vis += t2 - t0;
    }
    printf("kernel launch: call %.2f us, flag visible to host after %.2f us\n", call / reps, vis / reps);
  }
  {  // event record + wait
    cudaEvent_t ev;
    cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    double rec = 0, wait = 0;
    const int reps = 200;
    for (int r = 0; r < reps; r++) {
      const double t0 = now_us();
      cudaEventRecord(ev, s);
      const double t1 = now_us();
      cudaStreamWaitEvent(s2, ev, 0);
      const double t2 = now_us();
      rec += t1 - t0;
      wait += t2 - t1;
    }
    cudaDeviceSynchronize();
    printf("cudaEventRecord %.2f us, cudaStreamWaitEvent %.2f us\n", rec / reps, wait / reps);
  }
  {  // 10 back-to-back launches
    cudaDeviceSynchronize();
    const double t0 = now_us();
    for (int i = 0; i < 10; i++) tiny<<<1, 32, 0, s>>>(d_word);
    const double t1 = now_us();
    cudaDeviceSynchronize();
    printf("10 launches back to back: %.2f us of host time\n", t1 - t0);
  }
  return 0;
}
