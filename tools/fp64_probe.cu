// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/fp64_probe tools/fp64_probe.cu
// Micro-benchmark: issue cost of the instruction kinds the scan's inner loop
// uses (fp64 compare / multiply, 64-bit integer compare, REDUX), 8 warps / SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void probe(const double *in, unsigned long long *out, int iters, long long *cycles) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  double a0 = in[tid & 1023], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  double b = in[(tid + 7) & 1023];
  unsigned long long acc = 0, k0 = (unsigned long long)tid * 0x9E3779B97F4A7C15ull, kb = k0 >> 3;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (KIND == 0) {  // 8 independent fp64 compares
      acc += (a0 >= b) + (a1 >= b) + (a2 >= b) + (a3 >= b) + (a0 <= b) + (a1 <= b) + (a2 <= b) + (a3 <= b);
      b += 0.5;
    } else if (KIND == 1) {  // 4 fp64 multiplies + compares
      acc += (__dmul_rn(a0, b) >= a1) + (__dmul_rn(a1, b) >= a2) + (__dmul_rn(a2, b) >= a3) + (__dmul_rn(a3, b) >= a0);
      b += 0.5;
    } else if (KIND == 2) {  // 8 independent 64-bit integer compares
      acc += (k0 >= kb) + (k0 + 1 >= kb) + (k0 + 2 >= kb) + (k0 + 3 >= kb) + (k0 <= kb) + (k0 + 5 <= kb) + (k0 + 6 <= kb) + (k0 + 7 <= kb);
      kb += 0x1234567ull;
    } else if (KIND == 3) {  // warp reductions
      acc += __reduce_min_sync(0xFFFFFFFFu, (unsigned)(k0 >> 32) + i);
      acc += __reduce_min_sync(0xFFFFFFFFu, (unsigned)k0 + i);
    } else if (KIND == 4) {  // vote + ballot
      acc += __any_sync(0xFFFFFFFFu, ((k0 + i) & 7) == 0);
      acc += __ballot_sync(0xFFFFFFFFu, ((k0 + i) & 3) == 0);
    }
  }
  const long long t1 = clock64();
  out[tid] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  double *in; unsigned long long *out; long long *cyc;
  cudaMalloc(&in, 1024 * 8); cudaMemset(in, 0, 1024 * 8);
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  const char *names[] = {"8x DSETP", "4x DMUL+DSETP", "8x 64-bit ISETP", "2x REDUX.min", "VOTE.ANY + BALLOT"};
  for (int warps = 8; warps <= 32; warps *= 2) {
    for (int k = 0; k < 5; ++k) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (k) {
          case 0: probe<0><<<148, warps * 32>>>(in, out, iters, cyc); break;
          case 1: probe<1><<<148, warps * 32>>>(in, out, iters, cyc); break;
          case 2: probe<2><<<148, warps * 32>>>(in, out, iters, cyc); break;
          case 3: probe<3><<<148, warps * 32>>>(in, out, iters, cyc); break;
          case 4: probe<4><<<148, warps * 32>>>(in, out, iters, cyc); break;
        }
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      printf("%2d warps/SM  %-18s %.1f cycles per loop iteration per warp\n", warps, names[k], (double)h / iters);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
