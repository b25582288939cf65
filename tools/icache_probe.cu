// Does prefetching a kernel's own code (as data, into L2) hide cold
// instruction fetch? After the L2 flush between bench steps every instruction
// line of a kernel is a DRAM round trip the first time it is executed
// (profiles/round2_timeline.md); this probe times a long straight-line kernel
//   cold (L2 flushed), cold + prefetch.global.L2 over its own code, and warm.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/icache_probe tools/icache_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}

#define S1 v = fmaf(v, 1.0f + (__COUNTER__ % 4001) * 1.1920929e-7f, 0.25f);
#define S4 S1 S1 S1 S1
#define S16 S4 S4 S4 S4
#define S64 S16 S16 S16 S16
#define S256 S64 S64 S64 S64
#define S1024 S256 S256 S256 S256

__device__ __noinline__ float body(float v) {
  S1024 S1024 S1024 S1024
  return v;
}

typedef float (*body_fn)(float);
__device__ body_fn g_body = body;

__global__ void __launch_bounds__(32) big(float *x, int prefetch_bytes, unsigned long long code, unsigned long long *t, unsigned long long *peek) {
  if (peek && threadIdx.x == 0) {
    peek[0] = (unsigned long long)(void *)g_body;
  }
  if (prefetch_bytes) {
    const char *base = (const char *)code;
    for (int i = threadIdx.x * 128; i < prefetch_bytes; i += 32 * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + i));
    __syncwarp();
    __nanosleep(3000);
  }
  const unsigned long long t0 = gtime();
  float v = x[threadIdx.x];
  v = body(v);
  const unsigned long long t1 = gtime();
  x[threadIdx.x] = v;
  if (threadIdx.x == 0) { t[0] = t1 - t0; }
}

__global__ void reader(const unsigned long long *p, unsigned long long *out, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = p[i];
}

int main() {
  float *x; unsigned long long *t, *peek; char *flush;
  const size_t FL = 512ull << 20;
  CK(cudaMalloc(&x, 128)); CK(cudaMalloc(&t, 64)); CK(cudaMalloc(&peek, 4096)); CK(cudaMalloc(&flush, FL));
  CK(cudaMemset(x, 0, 128));
  unsigned long long h[8], ht;
  // 1. the device-side address of the kernel
  big<<<1, 32>>>(x, 0, 0, t, peek);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, peek, 8, cudaMemcpyDeviceToHost));
  printf("device-side &big = 0x%llx\n", h[0]);
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, big));
  printf("binaryVersion %d regs %d\n", fa.binaryVersion, fa.numRegs);
  // 2. can it be read as data?
  reader<<<1, 32>>>((const unsigned long long *)h[0], peek + 8, 16);
  cudaError_t e = cudaDeviceSynchronize();
  printf("read code as data: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 2;
  unsigned long long w[16];
  CK(cudaMemcpy(w, peek + 8, sizeof w, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 16; i += 2) printf("  +%3d: 0x%016llx 0x%016llx\n", i * 8, w[i], w[i + 1]);
  const int CODE = 64 * 1024;
  for (int rep = 0; rep < 3; ++rep) {
    for (int mode = 0; mode < 3; ++mode) {
      if (mode != 2) { CK(cudaMemset(flush, rep + mode, FL)); CK(cudaDeviceSynchronize()); }
      big<<<1, 32>>>(x, mode == 1 ? CODE : 0, h[0], t, nullptr);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(&ht, t, 8, cudaMemcpyDeviceToHost));
      printf("rep %d %-16s %8.2f us\n", rep, mode == 0 ? "cold" : mode == 1 ? "cold+prefetch" : "warm", ht / 1e3);
    }
  }
  return 0;
}
