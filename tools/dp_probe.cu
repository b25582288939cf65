// dp_probe.cu -- latencies that shape the chain-DP block (one warp, B200):
// dependent fp64 adds / compares, 64-bit shuffles, shared-memory atomics,
// REDUX, and three codings of the C-vector recurrence of skyopt_step.cuh.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/dp_probe tools/dp_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int N = 1024;

__device__ __forceinline__ uint64_t pkey(double p) {
  uint64_t b = (uint64_t)__double_as_longlong(p);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double kprice(uint64_t k) {
  uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double((long long)b);
}

__global__ void probe(double *out, long long *cyc, const double *in, int T, int C) {
  __shared__ double s_in[N];
  __shared__ unsigned long long s_at[64];
  __shared__ double s_D[32][32], s_tar[32][32];
  __shared__ unsigned long long s_mv[32][32];
  const int lane = threadIdx.x;
  for (int i = lane; i < N; i += 32) s_in[i] = in[i];
  for (int i = lane; i < 64; i += 32) s_at[i] = ~0ull;
  for (int i = lane; i < 32 * 32; i += 32) { s_tar[i / 32][i % 32] = in[i % N] * 0.5; s_mv[i / 32][i % 32] = pkey(in[(i * 7) % N]); }
  __syncwarp();
  long long t0, t1;
  int k = 0;
  // 1. dependent DADD
  { double x = in[lane]; t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __dadd_rn(x, 1.25);
    t1 = clock64(); out[k] = x; if (lane == 0) cyc[k] = t1 - t0; ++k; }
  // 2. dependent compare-select with independent operands (DSETP + FSEL)
  { double b = 1e300; t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { const double v = s_in[i]; if (v < b) b = v; }
    t1 = clock64(); out[k] = b; if (lane == 0) cyc[k] = t1 - t0; ++k; }
  // 3. fmin chain
  { double b = 1e300; t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) b = fmin(b, s_in[i]);
    t1 = clock64(); out[k] = b; if (lane == 0) cyc[k] = t1 - t0; ++k; }
  // 4. 64-bit shuffle chain
  { double x = in[lane]; t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xFFFFFFFFu, x, (lane + 1) & 31);
    t1 = clock64(); out[k] = x; if (lane == 0) cyc[k] = t1 - t0; ++k; }
  // 5. u64 key min chain (integer compare)
  { uint64_t b = ~0ull; t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { const uint64_t v = pkey(s_in[i]); if (v < b) b = v; }
    t1 = clock64(); out[k] = kprice(b); if (lane == 0) cyc[k] = t1 - t0; ++k; }
  // 6. REDUX chain
  { uint32_t x = lane; t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __reduce_min_sync(0xFFFFFFFFu, x + lane + i);
    t1 = clock64(); out[k] = x; if (lane == 0) cyc[k] = t1 - t0; ++k; }
  // 7. shared atomicMin u64, one address for the whole warp
  { t0 = clock64();
    for (int i = 0; i < 64; ++i) atomicMin(&s_at[0], (unsigned long long)(N * 64 - i * 32 - lane));
    __syncwarp(); t1 = clock64(); out[k] = (double)s_at[0]; if (lane == 0) cyc[k] = (t1 - t0) * (N / 64); ++k; }
  // 8. shared atomicMin u64, four addresses
  { t0 = clock64();
    for (int i = 0; i < 64; ++i) atomicMin(&s_at[1 + (lane & 3)], (unsigned long long)(N * 64 - i * 32 - lane));
    __syncwarp(); t1 = clock64(); out[k] = (double)s_at[1]; if (lane == 0) cyc[k] = (t1 - t0) * (N / 64); ++k; }
  // 9. DADD -> DADD -> key (what one candidate costs in the winners step), 8 independent
  { double acc = 0; t0 = clock64();
    for (int i = 0; i < N; i += 8) {
      double v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __dadd_rn(__dadd_rn(s_in[i + j], 3.5), 0.25);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += v[j];
    }
    t1 = clock64(); out[k] = acc; if (lane == 0) cyc[k] = t1 - t0; ++k; }
  // ---- recurrence codings, T steps over C clouds; cycles for all T steps (scaled to N/T)
  const double kInf = __longlong_as_double(0x7FF0000000000000ll);
  // 10. round-1 coding: D in shared memory, lane < C, __syncwarp per step
  { t0 = clock64();
    for (int lt = 0; lt < T; ++lt) {
      if (lane < C) {
        double b;
        if (lt == 0) b = s_tar[lt][lane];
        else { b = kInf; for (int g = 0; g < C; ++g) { const double e = (g != lane) ? s_tar[lt][g] : 0.0; const double v = __dadd_rn(s_D[lt - 1][g], e); if (v < b) b = v; } }
        s_D[lt][lane] = __dadd_rn(kprice(s_mv[lt][lane]), b);
      }
      __syncwarp();
    }
    t1 = clock64(); out[k] = s_D[T - 1][lane % C]; if (lane == 0) cyc[k] = t1 - t0; ++k; }
  // 11. round-2 coding (skyopt_step.cuh today): D in a register, shuffles
  { const int h = lane < C ? lane : 0; double dprev = 0.0; t0 = clock64();
    for (int lt = 0; lt < T; ++lt) {
      double b;
      if (lt == 0) b = s_tar[lt][h];
      else { b = kInf;
#pragma unroll 4
        for (int g = 0; g < C; ++g) { const double dg = __shfl_sync(0xFFFFFFFFu, dprev, g); const double e = (g != lane) ? s_tar[lt][g] : 0.0; const double v = __dadd_rn(dg, e); if (v < b) b = v; } }
      dprev = __dadd_rn(kprice(s_mv[lt][h]), b);
    }
    t1 = clock64(); out[k] = dprev; if (lane == 0) cyc[k] = t1 - t0; ++k; }
  // 12. lane = (h, g) pair, keys, butterfly min inside each group of Cp lanes
  { int Cp = 1; while (Cp < C) Cp <<= 1;
    const int h = lane / Cp, g = lane % Cp; const bool on = h < C && g < C && h * Cp + g < 32;
    double dmine = 0.0;  // lane (h, g) holds D[t-1][g]
    t0 = clock64();
    for (int lt = 0; lt < T; ++lt) {
      uint64_t key = ~0ull;
      if (on) {
        if (lt == 0) { if (g == 0) key = pkey(s_tar[lt][h]); }
        else { const double e = (g != h) ? s_tar[lt][g] : 0.0; key = pkey(__dadd_rn(dmine, e)); }
      }
      for (int o = Cp >> 1; o; o >>= 1) { const uint64_t ok = __shfl_xor_sync(0xFFFFFFFFu, key, o); if (ok < key) key = ok; }
      // every lane of group h now has B[h]; D[h] = mv[h] + B[h]
      const double dh = on ? __dadd_rn(kprice(s_mv[lt][h]), kprice(key)) : 0.0;
      // lane (h, g) needs D[g]: it sits in lane (g, *)
      dmine = __shfl_sync(0xFFFFFFFFu, dh, (g < C ? g : 0) * Cp);
    }
    t1 = clock64(); out[k] = dmine; if (lane == 0) cyc[k] = t1 - t0; ++k; }
  // 13. empty clock pair
  { t0 = clock64(); t1 = clock64(); out[k] = 0; if (lane == 0) cyc[k] = t1 - t0; ++k; }
}

int main() {
  double *in, *out; long long *cyc;
  cudaMallocManaged(&in, N * 8); cudaMallocManaged(&out, 64 * 8); cudaMallocManaged(&cyc, 64 * 8);
  for (int i = 0; i < N; ++i) in[i] = 1.0 + ((i * 2654435761u) % 1000) / 997.0;
  const char *names[] = {"dependent DADD", "DSETP+FSEL min chain", "fmin chain", "shfl 64-bit chain", "u64 key min chain",
                         "REDUX chain", "ATOMS min u64, 1 address (x32 lanes)", "ATOMS min u64, 4 addresses",
                         "2 DADD per candidate, 8-way ILP", "recurrence: smem D (round 1)", "recurrence: shfl D (round 2)",
                         "recurrence: (h,g) lanes + key butterfly", "clock pair"};
  for (int rep = 0; rep < 2; ++rep) {
    probe<<<1, 32>>>(out, cyc, in, 32, 4);
    cudaDeviceSynchronize();
  }
  printf("%-45s %12s %10s\n", "probe", "cycles", "per op");
  for (int k = 0; k < 13; ++k) {
    const double per = (k >= 9 && k <= 11) ? (double)cyc[k] / 32 : (double)cyc[k] / N;
    printf("%-45s %12lld %10.1f%s\n", names[k], cyc[k], per, (k >= 9 && k <= 11) ? " /step (T=32,C=4)" : "");
  }
  probe<<<1, 32>>>(out, cyc, in, 32, 8); cudaDeviceSynchronize();
  for (int k = 9; k <= 11; ++k) printf("C=8  %-40s %12lld %10.1f /step\n", names[k], cyc[k], (double)cyc[k] / 32);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
