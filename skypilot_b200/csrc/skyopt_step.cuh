// skyopt_step.cuh -- the whole optimizer step as ONE persistent cooperative
// kernel: scan2 (filter + argmin) -> grid barrier -> place (expansion, blocked
// filter, cost; block per task) -> chain DP by the block that finishes a DAG's
// last task. One launch instead of five: on a 30 us step the launch gaps and
// cold prologues of separate kernels were half of the time
// (profiles/round2_timeline.md).
#pragma once

#include "skyopt_fast.cuh"

namespace skyopt {

// Chain DP (reference sky/optimizer.py:429-487) of one DAG by one block.
//
// The reference walks the chain with, per child candidate c and parent
// candidate p, the sum fl(dp[p] + egress(cloud(p), cloud(c))) and keeps the
// FIRST minimum over p. Egress only depends on the two clouds, so with
//   vmin[t][g] = smallest candidate value of task t in cloud g, idx[t][g] the
//                first candidate holding it (left by the task's place block),
//   D[t][g]    = fl(vmin[t][g] + B[t][g])     (dp of that candidate),
//   B[t][h]    = min_g fl(D[t-1][g] + e_t(g, h)),  e_t(g, h) = tar_t[g] (g != h), 0 (g == h)
// the recurrence is C wide, and the argmin over g -- ties to the smaller
// candidate index, i.e. the first minimum in candidate order -- IS the
// reference's best parent of every child candidate in cloud h: fl() is
// monotone, so within a cloud no candidate beats the one with the smallest
// value. Same sums, same minima as optimizer.py:456-470, and no candidate is
// ever read. One case needs the full tables: a candidate in FRONT of idx[t][g]
// with a larger value whose sum rounds to the same double would be the
// reference's first minimum. v2[t][g] (the smallest value in front) makes
// that checkable: if fl(fl(v2 + B) + e) == fl(fl(vmin + B) + e) for any cloud
// the block falls back to chain_full() below, which evaluates every candidate.
struct ChainSmem {
  double tar[kFastTasks][SKYOPT_MAX_CLOUDS];
  unsigned long long mv[kFastTasks][SKYOPT_MAX_CLOUDS];  // price_key(vmin)
  unsigned long long v2[kFastTasks][SKYOPT_MAX_CLOUDS];  // price_key(v2)
  int mi[kFastTasks][SKYOPT_MAX_CLOUDS];                 // (idx << 5) | cloud
  double B[kFastTasks][SKYOPT_MAX_CLOUDS];
  long long toff[kFastTasks];
  int bk[kFastTasks + 1][SKYOPT_MAX_CLOUDS];  // best parent (idx << 5 | cloud) by child task, child cloud
  int tn[kFastTasks], np[kFastTasks], src[kFastTasks], choice[kFastTasks];
  double obj;
  int first_empty, hazard;
};

__device__ __forceinline__ void step_mark(unsigned long long *trace, int slot) {
  if (trace && threadIdx.x == 0 && blockIdx.x < kTraceBlocks)
    trace[((size_t)2 * kTraceBlocks + blockIdx.x) * kTraceSlots + slot] = global_ns();
}

// (price key, id) pairs in lexicographic order, branch-free
constexpr unsigned long long kKeyInf = 0xFFF0000000000000ull;  // price_key(+inf)
__device__ __forceinline__ bool key_less(unsigned long long ak, int ai, unsigned long long bk, int bi) {
  return (ak < bk) | ((ak == bk) & (ai < bi));
}

// The full evaluation: winners from every candidate's own sum, candidates read
// from the task tables in global memory. Only runs after a hazard (see above),
// so it is written for size, not speed.
__device__ __noinline__ void chain_full(ChainSmem &M, const SolveWork &w, int T, int C, int Cp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double kInf = __longlong_as_double(0x7FF0000000000000ll);
  const int parts = 32 / Cp;
  const int h = lane % Cp, part = lane / Cp;
#pragma unroll 1
  for (int lt = warp; lt < T; lt += kFastWarps) {
    const bool sink = lt == T - 1;
    const int n = M.tn[lt];
    const long long toff = M.toff[lt];
    double bv = kInf; int bi = 0x7FFFFFFF;
    const bool wanted = h < C && (sink ? h == 0 : M.mv[lt + 1][h] != kKeyNone);
    if (wanted) {
#pragma unroll 1
      for (int p = part; p < n; p += parts) {
        const int cp = __ldcg(w.tc_cloud + toff + p) & (SKYOPT_MAX_CLOUDS - 1);
        const double dpp = __dadd_rn(__ldcg(w.tc_value + toff + p), M.B[lt][cp]);
        const double e = (!sink && cp != h) ? M.tar[lt + 1][cp] : 0.0;
        const double sum = sink ? dpp : __dadd_rn(dpp, e);
        if (sum < bv) { bv = sum; bi = p; }
      }
    }
#pragma unroll 1
    for (int o = Cp; o < 32; o <<= 1) {
      const double ov = __shfl_xor_sync(0xFFFFFFFFu, bv, o);
      const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
      lexmin(bv, bi, ov, oi);
    }
    if (part == 0 && h < C && (!sink || h == 0)) {
      int cl = 0;
      if (bi != 0x7FFFFFFF && bi < n) cl = __ldcg(w.tc_cloud + toff + bi) & (SKYOPT_MAX_CLOUDS - 1);
      M.bk[lt + 1][h] = (int)(((unsigned)bi << 5) | (unsigned)cl);
      if (sink) M.obj = bv;
    }
  }
}

// The recurrence for any number of clouds and any sign of the values: price
// keys (a total order on doubles), butterfly minima for many clouds.
__device__ __noinline__ void recurrence_general(ChainSmem &M, int T, int C, int Cp, bool &hz, double &bv, int &bi) {
  const int lane = threadIdx.x & 31;
  const double kInf = __longlong_as_double(0x7FF0000000000000ll);
  const int h = lane < C ? lane : 0;
  const bool live = lane < C;
  double dprev = 0.0, hprev = 0.0;   // D[t-1][lane], fl(v2[t-1][lane] + B[t-1][lane])
  int myid = 0x7FFFFFFF;
  bool has2 = false;
  // the next task's operands are fetched one step ahead of the dependent chain
  double tar_n = M.tar[0][h];
  unsigned long long mk_n = M.mv[0][h], v2_n = M.v2[0][h];
  int mi_n = M.mi[0][h], np_n = M.np[0];
#pragma unroll 1
  for (int lt = 0; lt < T; ++lt) {
    const double tar = tar_n;
    const unsigned long long mk = mk_n, v2k = v2_n;
    const int mi = mi_n, np = np_n;
    if (lt + 1 < T) {
      tar_n = M.tar[lt + 1][h]; mk_n = M.mv[lt + 1][h]; v2_n = M.v2[lt + 1][h];
      mi_n = M.mi[lt + 1][h]; np_n = M.np[lt + 1];
    }
    double b;
    if (np == 0) {
      b = tar;  // dummy source: 0 + egress from the inputs' cloud
    } else {
      const double mine = live ? __dadd_rn(dprev, tar) : kInf;
      // a candidate in front of the cheapest one whose sum rounds the same
      if (live && has2 && dprev < kInf && (hprev == dprev || __dadd_rn(hprev, tar) == mine)) hz = true;
      // The minima are taken on price keys: a total order on doubles as
      // 64-bit integers (equal doubles <=> equal keys for the non-negative
      // sums here). An fp64 compare-and-select costs four times the latency
      // of the integer one, and this loop is nothing but dependent compares.
      const unsigned long long kd = price_key(dprev);
      const unsigned long long km = price_key(mine);
      const int idm = live ? myid : 0x7FFFFFFF;
      unsigned long long bkey = kd;   // own cloud: egress 0
      int bid = myid;
      if (Cp <= 8) {
        // few clouds: every lane reads every cloud's sum (independent
        // shuffles) and keeps a running first minimum -- a third of the
        // instructions of the butterfly below, and one warp's time here is
        // its dependent-instruction count
#pragma unroll 4
        for (int g = 0; g < C; ++g) {
          const unsigned long long kg = __shfl_sync(0xFFFFFFFFu, km, g);
          const int ig = __shfl_sync(0xFFFFFFFFu, idm, g);
          const bool take = (g != lane) & key_less(kg, ig, bkey, bid);
          bkey = take ? kg : bkey;
          bid = take ? ig : bid;
        }
      } else {
        // smallest and second smallest sum (with their ids) by a butterfly;
        // a lane whose own sum is the smallest takes the second
        unsigned long long k1 = km, k2 = kKeyInf;
        int i1 = idm, i2 = 0x7FFFFFFF;
#pragma unroll 1
        for (int o = 1; o < Cp; o <<= 1) {
          const unsigned long long p1 = __shfl_xor_sync(0xFFFFFFFFu, k1, o);
          const int q1 = __shfl_xor_sync(0xFFFFFFFFu, i1, o);
          const unsigned long long p2 = __shfl_xor_sync(0xFFFFFFFFu, k2, o);
          const int q2 = __shfl_xor_sync(0xFFFFFFFFu, i2, o);
          const bool pl = key_less(p1, q1, k1, i1);
          const unsigned long long lo = pl ? p1 : k1, hi = pl ? k1 : p1;
          const int loi = pl ? q1 : i1, hii = pl ? i1 : q1;
          const bool sl = key_less(p2, q2, k2, i2);
          const unsigned long long s2 = sl ? p2 : k2;
          const int s2i = sl ? q2 : i2;
          const bool tl = key_less(s2, s2i, hi, hii);
          k1 = lo; i1 = loi;
          k2 = tl ? s2 : hi; i2 = tl ? s2i : hii;
        }
        const bool own = i1 == myid;   // ids are distinct: (candidate index, cloud)
        const unsigned long long ok = own ? k2 : k1;
        const int oi = own ? i2 : i1;
        const bool ol = key_less(ok, oi, kd, myid);
        bkey = ol ? ok : kd;
        bid = ol ? oi : myid;
      }
      b = key_price(bkey);
      if (live) M.bk[lt][lane] = bid;
    }
    if (live) M.B[lt][lane] = b;
    dprev = (mk == kKeyNone || !live) ? kInf : __dadd_rn(key_price(mk), b);
    has2 = v2k != kKeyNone;
    hprev = has2 ? __dadd_rn(key_price(v2k), b) : 0.0;
    myid = mi;
  }
  // the sink: egress 0 from every cloud, first minimum of D[T-1][.]
  if (live && has2 && dprev < kInf && hprev == dprev) hz = true;
  bv = live ? dprev : kInf;
  bi = live ? myid : 0x7FFFFFFF;
#pragma unroll 1
  for (int o = 1; o < Cp; o <<= 1) {
    const double ov = __shfl_xor_sync(0xFFFFFFFFu, bv, o);
    const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
    lexmin(bv, bi, ov, oi);
  }
}

// shared memory by 32-bit address: the loop below must not re-derive the
// shared window of a generic pointer at every access
__device__ __forceinline__ unsigned long long lds_u64(unsigned a) {
  unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v;
}
__device__ __forceinline__ int lds_s32(unsigned a) {
  int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v;
}
__device__ __forceinline__ void sts_u64(unsigned a, unsigned long long v) {
  asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
__device__ __forceinline__ void sts_s32(unsigned a, int v) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void lexmin_bits(unsigned long long &k, int &i, unsigned long long ok, int oi) {
  const bool lt = (ok < k) | ((ok == k) & (oi < i));
  k = lt ? ok : k; i = lt ? oi : i;
}

// The same recurrence for at most four clouds and non-negative values (every
// real request): one warp's time is its chain of dependent instructions, so
// this loop is written for the shortest chain -- sums compared as raw bit
// patterns (monotone for non-negative doubles), every lane reads the other
// clouds' sums with independent shuffles and reduces them in a tree, operands
// prefetched one step ahead through 32-bit shared addresses.
template <int CP>
__device__ __forceinline__ void recurrence_lean(ChainSmem &M, int T, int C, bool &hz, double &bv, int &bi) {
  const int lane = threadIdx.x & 31;
  const bool live = lane < C;
  const int h = live ? lane : 0;
  constexpr unsigned long long kInfBits = 0x7FF0000000000000ull;
  constexpr unsigned long long kTop = 0x8000000000000000ull;
  unsigned a_tar = (unsigned)__cvta_generic_to_shared(&M.tar[0][h]);
  unsigned a_mv = (unsigned)__cvta_generic_to_shared(&M.mv[0][h]);
  unsigned a_v2 = (unsigned)__cvta_generic_to_shared(&M.v2[0][h]);
  unsigned a_mi = (unsigned)__cvta_generic_to_shared(&M.mi[0][h]);
  unsigned a_np = (unsigned)__cvta_generic_to_shared(&M.np[0]);
  unsigned a_B = (unsigned)__cvta_generic_to_shared(&M.B[0][h]);
  unsigned a_bk = (unsigned)__cvta_generic_to_shared(&M.bk[0][h]);
  constexpr unsigned kRow8 = SKYOPT_MAX_CLOUDS * 8, kRow4 = SKYOPT_MAX_CLOUDS * 4;
  unsigned long long d = 0, h2 = 0;   // bits of D[t-1][lane], of fl(v2[t-1][lane] + B[t-1][lane])
  int myid = 0x7FFFFFFF;
  bool has2 = false;
  unsigned long long tar_n = lds_u64(a_tar), mk_n = lds_u64(a_mv), v2_n = lds_u64(a_v2);
  int mi_n = lds_s32(a_mi), np_n = lds_s32(a_np);
#pragma unroll 1
  for (int lt = 0; lt < T; ++lt) {
    const unsigned long long tar = tar_n, mk = live ? mk_n : kKeyNone, v2k = v2_n;
    const int mi = live ? mi_n : 0x7FFFFFFF, np = np_n;
    a_tar += kRow8; a_mv += kRow8; a_v2 += kRow8; a_mi += kRow4; a_np += 4;
    if (lt + 1 < T) {
      tar_n = lds_u64(a_tar); mk_n = lds_u64(a_mv); v2_n = lds_u64(a_v2);
      mi_n = lds_s32(a_mi); np_n = lds_s32(a_np);
    }
    unsigned long long b;
    if (np == 0) {
      b = tar;  // dummy source: 0 + egress from the inputs' cloud
    } else {
      const double dd = __longlong_as_double((long long)d), td = __longlong_as_double((long long)tar);
      const unsigned long long ka = (unsigned long long)__double_as_longlong(__dadd_rn(dd, td));
      if (has2 && d < kInfBits &&
          (h2 == d || (unsigned long long)__double_as_longlong(
                          __dadd_rn(__longlong_as_double((long long)h2), td)) == ka))
        hz = true;
      unsigned long long e[CP]; int ei[CP];
#pragma unroll
      for (int g = 0; g < CP; ++g) {
        const unsigned long long kg = __shfl_sync(0xFFFFFFFFu, ka, g);
        const int ig = __shfl_sync(0xFFFFFFFFu, myid, g);
        const bool own = g == lane;          // own cloud: egress 0
        e[g] = own ? d : kg; ei[g] = ig;
      }
#pragma unroll
      for (int w = 1; w < CP; w <<= 1)
#pragma unroll
        for (int g = 0; g + w < CP; g += 2 * w) lexmin_bits(e[g], ei[g], e[g + w], ei[g + w]);
      b = e[0];
      if (live) sts_s32(a_bk, ei[0]);
    }
    if (live) sts_u64(a_B, b);
    a_B += kRow8; a_bk += kRow4;
    const double bd = __longlong_as_double((long long)b);
    d = (mk == kKeyNone) ? kInfBits
                         : (unsigned long long)__double_as_longlong(
                               __dadd_rn(__longlong_as_double((long long)(mk ^ kTop)), bd));
    has2 = v2k != kKeyNone;
    h2 = has2 ? (unsigned long long)__double_as_longlong(
                    __dadd_rn(__longlong_as_double((long long)(v2k ^ kTop)), bd))
              : 0ull;
    myid = mi;
  }
  // the sink: egress 0 from every cloud, first minimum of D[T-1][.]
  if (has2 && d < kInfBits && h2 == d) hz = true;
  unsigned long long k = d; int id = myid;
#pragma unroll
  for (int o = 1; o < CP; o <<= 1) {
    const unsigned long long ok = __shfl_xor_sync(0xFFFFFFFFu, k, o);
    const int oi = __shfl_xor_sync(0xFFFFFFFFu, id, o);
    lexmin_bits(k, id, ok, oi);
  }
  bv = __longlong_as_double((long long)k);
  bi = id;
}

__device__ __forceinline__ void chain_body(const CatDev &cat, const SolveIn &in, const SolveWork &w,
                                           const SolveOut &out, const TaskMin *task_mv,
                                           int dag, int force_full, unsigned char *smem) {
  ChainSmem &M = *reinterpret_cast<ChainSmem *>(smem);
  const SkyoptDag D = in.dags[dag];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = cat.n_clouds;
  const int T = D.task_end - D.task_begin;
  const double kInf = __longlong_as_double(0x7FF0000000000000ll);
  step_mark(out.trace, 0);
  if (tid == 0) { M.first_empty = 0x7FFFFFFF; M.hazard = force_full & 1; }
  int signed_values = (force_full & 2) ? 1 : 0;  // a negative value or tariff anywhere in the DAG
  if (tid < T) {
    const int t = D.task_begin + tid;
    const SkyoptTask TK = in.tasks[t];
    M.tn[tid] = __ldcg(out.task_n + t);
    M.toff[tid] = in.task_off[t];
    M.np[tid] = TK.n_parents;
    M.src[tid] = TK.n_parents ? TK.edge_tariff_begin : TK.src_tariff_begin;
  }
#pragma unroll 1
  for (int i = tid; i < T * C; i += kScanThreads) {
    const int lt = i / C, cc = i % C;
    // the cheapest candidate of (task, cloud): left by the task's place block
    const ulonglong2 *pm = reinterpret_cast<const ulonglong2 *>(task_mv + (int64_t)(D.task_begin + lt) * C + cc);
    const ulonglong2 m0 = __ldcg(pm), m1 = __ldcg(pm + 1);
    M.mv[lt][cc] = m0.x;
    M.v2[lt][cc] = m0.y;
    // price keys of non-negative doubles have the top bit set
    if (!(m0.x >> 63) || !(m0.y >> 63)) signed_values = 1;
    M.mi[lt][cc] = (int)(((unsigned)m1.x << 5) | (unsigned)cc);
  }
  __syncthreads();
  if (tid < T && M.tn[tid] == 0) atomicMin(&M.first_empty, tid);
#pragma unroll 1
  for (int i = tid; i < T * C; i += kScanThreads) {
    const int lt = i / C, cc = i % C;
    const double tv = M.src[lt] >= 0 ? in.tariffs[M.src[lt] + cc] : 0.0;
    M.tar[lt][cc] = tv;
    if (__double_as_longlong(tv) < 0 || tv != tv) signed_values = 1;
  }
  const bool lean = __syncthreads_or(signed_values) == 0;
  if (M.first_empty != 0x7FFFFFFF) {
    if (tid == 0) {
      SkyoptDagResult r; r.status = 1; r.task_fail = M.first_empty;
      r.objective = __longlong_as_double(0x7FF8000000000000ll);
      out.dag[dag] = r;
    }
    return;
  }
  step_mark(out.trace, 1);
  const long long c_start = clock64();
  int Cp = 1;
  while (Cp < C) Cp <<= 1;
  if (warp == 0) {
    bool hz = false;
    double bv; int bi;
    if (lean && Cp <= 4) {
      if (Cp == 4) recurrence_lean<4>(M, T, C, hz, bv, bi);
      else if (Cp == 2) recurrence_lean<2>(M, T, C, hz, bv, bi);
      else recurrence_lean<1>(M, T, C, hz, bv, bi);
    } else {
      recurrence_general(M, T, C, Cp, hz, bv, bi);
    }
    const bool any_hz = __any_sync(0xFFFFFFFFu, hz);
    if (lane == 0) {
      M.bk[T][0] = bi; M.obj = bv;
      if (any_hz) M.hazard = 1;
      if (out.trace && blockIdx.x < kTraceBlocks)
        out.trace[((size_t)2 * kTraceBlocks + blockIdx.x) * kTraceSlots + 8] = (unsigned long long)(clock64() - c_start);
    }
  }
  __syncthreads();
  step_mark(out.trace, 3);
  if (M.hazard) {
    chain_full(M, w, T, C, Cp);
    __syncthreads();
  }
  step_mark(out.trace, 4);
  if (tid == 0) {
    int id = M.bk[T][0];
#pragma unroll 1
    for (int lt = T - 1; lt >= 0; --lt) {
      M.choice[lt] = (int)((unsigned)id >> 5);
      if (lt > 0) id = M.bk[lt][id & (SKYOPT_MAX_CLOUDS - 1)];
    }
    SkyoptDagResult r; r.status = 0; r.task_fail = -1; r.objective = M.obj;
    out.dag[dag] = r;
  }
  __syncthreads();
  step_mark(out.trace, 5);
  if (tid < T) {
    const int t = D.task_begin + tid;
    const int ci = M.choice[tid];
    out.chosen_index[t] = ci;
    if (!in.tables_only) {
      const int64_t o = M.toff[tid] + ci;
      const int64_t ref = __ldcg(w.tc_ref + o);
      const int s = __ldcg(w.tc_slot + o);
      SkyoptCandidate c;
      c.slot = s;
      c.inst_id = __ldcg(in.ex.slot_inst + s);
      c.region_id = __ldcg(in.ex.cand_region + ref);
      c.zone_id = __ldcg(in.ex.cand_zone + ref);
      c.hourly = __ldcg(w.tc_hourly + o);
      c.value = __ldcg(w.tc_value + o);
      out.chosen[t] = c;
    }
  }
  step_mark(out.trace, 6);
}

struct StepArgs {
  Scan2Args scan;
  PlaceArgs place;
  SolveOut out;
  const int32_t *task_dag;
  int n_tasks, n_queries;
  const char *in_base;    // the uploaded input region (descriptors), prefetched into L2
  int64_t in_lines;
  int do_solve;           // every DAG is a chain of <= kFastTasks tasks
  int force_full;         // test knobs: bit 0 chain_full() on every DAG (SKYOPT_EXP=8), bit 1 the general recurrence (SKYOPT_EXP=16)
  int32_t *dag_done;      // [n_dags] tasks placed so far (zero between launches)
  unsigned int *sync;     // arrivals at the grid barrier, counted across launches
  unsigned int sync_target;  // ... and the count that completes this launch's barrier
  // the other copy of the scan results: re-armed here for the next launch of a
  // device-resident loop, by blocks that have passed the barrier
  uint32_t *next_best_rank, *next_any1;
  unsigned int *next_group_ready;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kScanThreads, kScanBlocksPerSM) step_kernel(StepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_step[];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  if (blockIdx.x == 0 && tid == 0 && a.scan.zero_flag) *a.scan.zero_flag = 0;
  // the problem's descriptors (tens of KB) are wanted in L2 by the later
  // phases: every block touches its share of the lines now
  for (int64_t line = (int64_t)blockIdx.x * kScanThreads + tid; line < a.in_lines;
       line += (int64_t)gridDim.x * kScanThreads)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.in_base + line * 128));
  if (a.scan.n_pieces) scan2_body(a.scan, smem_step);
  // ---- grid barrier: every scan result is published (cooperative launch:
  // all blocks are resident)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    atomicAdd(a.sync, 1u);
    // (the counter only grows: the host passes the total that ends this
    // launch's barrier, so nothing has to be reset on the way out)
    while ((int)(ld_acquire_u32(a.sync) - a.sync_target) < 0) if (!(a.scan.noprune & 4u)) __nanosleep(64);
    __threadfence();
  }
  __syncthreads();
  // ---- re-arm the other copy of the scan results (nobody reads it in this
  // launch): a few stores per block, off the critical path
#pragma unroll 1
  for (int i = blockIdx.x * kScanThreads + tid; i < a.n_queries; i += gridDim.x * kScanThreads) {
    a.next_best_rank[i] = kRankNone; a.next_any1[i] = 0u;
  }
  if (a.scan.group_ready)
#pragma unroll 1
    for (int i = blockIdx.x * kScanThreads + tid; i < a.scan.n_groups; i += gridDim.x * kScanThreads)
      a.next_group_ready[i] = 0u;
  for (int t = blockIdx.x; t < a.n_tasks; t += gridDim.x) {
    // which DAG, and how many tasks it has: asked for before the placement,
    // needed right after it
    int dag = 0, dag_tasks = 0;
    if (a.do_solve) {
      dag = a.task_dag[t];
      const SkyoptDag D = a.place.in.dags[dag];
      dag_tasks = D.task_end - D.task_begin;
    }
    place_body(a.place, t, smem_step);
    if (a.do_solve) {
      __syncthreads();  // the task's tables are written
      if (tid == 0) {
        __threadfence();
        const int done = atomicAdd(a.dag_done + dag, 1);
        s_last = (done == dag_tasks - 1) ? 1 : 0;
        if (s_last) a.dag_done[dag] = 0;  // ready for the next launch
      }
      __syncthreads();
      if (s_last) {
        __threadfence();
        chain_body(a.scan.cat, a.place.in, a.place.w, a.out, a.place.task_mv, dag, a.force_full, smem_step);
      }
    }
    __syncthreads();  // shared memory is reused by the next task
  }
}

}  // namespace skyopt
