// skyopt_step.cuh -- the whole optimizer step as ONE persistent cooperative
// kernel: scan2 (filter + argmin) -> grid barrier -> place (expansion, blocked
// filter, cost; block per task) -> chain DP by the block that finishes a DAG's
// last task. One launch instead of five: on a 30 us step the launch gaps and
// cold prologues of separate kernels were half of the time
// (profiles/round2_timeline.md).
#pragma once

#include "skyopt_fast.cuh"

namespace skyopt {

constexpr int kStepMaxCand = 4096;  // candidates of one DAG kept in shared memory

// Chain DP (reference sky/optimizer.py:429-487) of one DAG by one block of
// kScanThreads threads: the algorithm of solve_kernel's fast path -- per-cloud
// minima, a C-vector recurrence, then the first-minimum winners from exactly
// the sums the reference forms -- with the recurrence kept in registers.
struct ChainSmem {
  double tar[kFastTasks][SKYOPT_MAX_CLOUDS];
  unsigned long long mv[kFastTasks][SKYOPT_MAX_CLOUDS];  // price_key(min value)
  double B[kFastTasks][SKYOPT_MAX_CLOUDS];
  double cval[kStepMaxCand];
  long long toff[kFastTasks];
  int bk_idx[kFastTasks + 1][SKYOPT_MAX_CLOUDS];
  int tn[kFastTasks], np[kFastTasks], src[kFastTasks], cbase[kFastTasks + 1], choice[kFastTasks];
  unsigned char bk_cl[kFastTasks + 1][SKYOPT_MAX_CLOUDS];
  unsigned char ccl[kStepMaxCand];
  double obj;
  int first_empty, staged;
};

__device__ __forceinline__ void step_mark(unsigned long long *trace, int slot) {
  if (trace && threadIdx.x == 0 && blockIdx.x < kTraceBlocks)
    trace[((size_t)2 * kTraceBlocks + blockIdx.x) * kTraceSlots + slot] = global_ns();
}

__device__ __forceinline__ void chain_body(const CatDev &cat, const SolveIn &in, const SolveWork &w,
                                           const SolveOut &out, const unsigned long long *task_mv,
                                           int dag, unsigned char *smem) {
  ChainSmem &M = *reinterpret_cast<ChainSmem *>(smem);
  const SkyoptDag D = in.dags[dag];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = cat.n_clouds;
  const int T = D.task_end - D.task_begin;
  const double kInf = __longlong_as_double(0x7FF0000000000000ll);
  step_mark(out.trace, 0);
  if (tid == 0) M.first_empty = 0x7FFFFFFF;
  if (tid < T) {
    const int t = D.task_begin + tid;
    const SkyoptTask TK = in.tasks[t];
    M.tn[tid] = __ldcg(out.task_n + t);
    M.toff[tid] = in.task_off[t];
    M.np[tid] = TK.n_parents;
    M.src[tid] = TK.n_parents ? TK.edge_tariff_begin : TK.src_tariff_begin;
  }
  __syncthreads();
  if (tid < T && M.tn[tid] == 0) atomicMin(&M.first_empty, tid);
  if (tid == 0) {
    int acc = 0;
#pragma unroll 1
    for (int i = 0; i < T; ++i) { M.cbase[i] = acc; acc += M.tn[i]; }
    M.cbase[T] = acc;
    M.staged = acc <= kStepMaxCand ? 1 : 0;
  }
#pragma unroll 1
  for (int i = tid; i < T * C; i += kScanThreads) {
    const int lt = i / C, cc = i % C;
    // per-(task, cloud) minimum value: left by the task's place block
    M.mv[lt][cc] = __ldcg(task_mv + (int64_t)(D.task_begin + lt) * C + cc);
    M.tar[lt][cc] = M.src[lt] >= 0 ? in.tariffs[M.src[lt] + cc] : 0.0;
  }
  __syncthreads();
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
  const bool staged = M.staged != 0;
  int Cp = 1;
  while (Cp < C) Cp <<= 1;
  const int parts = 32 / Cp;
  // Two passes over ONE copy of the winners / back-tracking code. Pass 0: warp
  // 0 walks the recurrence (serial, one warp) while the other warps bring the
  // candidates to shared memory and then run the winners code on whatever B
  // holds -- a rehearsal whose only purpose is to have that code in the
  // instruction caches: this block is the only one that ever executes it, and
  // after a cold start every new instruction line is a DRAM round trip
  // (profiles/round2_timeline.md). Pass 1 is the real thing.
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 0) {
      if (warp == 0) {
        // B[t][h] = min_g fl(D[t-1][g] + e_t(g, h)),  D[t][h] = fl(mv[t][h] + B[t][h])
        // with e_t(g, h) = tar_t[g] for g != h and 0 for g == h. So
        //   B[t][h] = min(D[t-1][h], min_{g != h} A[g]),  A[g] = fl(D[t-1][g] + tar_t[g]):
        // lane g forms its one sum, a butterfly over the Cp lanes gives every
        // lane the smallest and second smallest A (with multiplicity), and
        // lane h takes the second one when its own A is the smallest. Same
        // sums, same minima as the reference's double loop
        // (optimizer.py:456-470), a third of the dependent latency.
        const int h = lane < C ? lane : 0;
        double dprev = 0.0;
#pragma unroll 1
        for (int lt = 0; lt < T; ++lt) {
          double b;
          if (M.np[lt] == 0) {
            b = M.tar[lt][h];  // dummy source: 0 + egress from the inputs' cloud
          } else {
            const double mine = lane < C ? __dadd_rn(dprev, M.tar[lt][h]) : kInf;
            double m1 = mine, m2 = kInf;
#pragma unroll 1
            for (int o = 1; o < Cp; o <<= 1) {
              const double p1 = __shfl_xor_sync(0xFFFFFFFFu, m1, o);
              const double p2 = __shfl_xor_sync(0xFFFFFFFFu, m2, o);
              const double lo = p1 < m1 ? p1 : m1, hi = p1 < m1 ? m1 : p1;
              const double s2 = p2 < m2 ? p2 : m2;
              m1 = lo; m2 = s2 < hi ? s2 : hi;
            }
            const double others = (mine == m1) ? m2 : m1;
            b = others < dprev ? others : dprev;
          }
          const unsigned long long mk = M.mv[lt][h];
          if (lane < C) M.B[lt][lane] = b;
          dprev = (mk == kKeyNone) ? kInf : __dadd_rn(key_price(mk), b);
        }
        if (out.trace && lane == 0 && blockIdx.x < kTraceBlocks)
          out.trace[((size_t)2 * kTraceBlocks + blockIdx.x) * kTraceSlots + 8] = (unsigned long long)(clock64() - c_start);
      } else if (staged) {
#pragma unroll 1
        for (int lt = warp - 1; lt < T; lt += kFastWarps - 1) {
          const int n = M.tn[lt];
          const long long toff = M.toff[lt];
          const int cb = M.cbase[lt];
#pragma unroll 1
          for (int c = lane; c < n; c += 32) {
            M.ccl[cb + c] = (unsigned char)__ldcg(w.tc_cloud + toff + c);
            M.cval[cb + c] = __ldcg(w.tc_value + toff + c);
          }
        }
      }
    }
    if (pass == 1 || warp != 0) {
      // winners: for parent task lt and child cloud h, the first minimum over
      // the parent's candidates p of fl(dp[p] + e_{lt+1}(cloud(p), h)) --
      // exactly the sums the reference forms (optimizer.py:456-470). A warp
      // takes a parent task; lane = (part, h): the candidates are dealt to
      // 32 / Cp parts, each lane walks its part in candidate order (strict '<'
      // keeps the first minimum), the parts are merged with (value, index)
      // comparisons. The last task's only child is the dummy sink (egress 0).
      const int h = lane % Cp, part = lane / Cp;
#pragma unroll 1
      for (int lt = warp; lt < T; lt += kFastWarps) {
        const bool sink = lt == T - 1;
        const int n = M.tn[lt];
        const long long toff = M.toff[lt];
        const int cb = M.cbase[lt];
        double bv = kInf; int bi = 0x7FFFFFFF;
        const bool wanted = h < C && (sink ? h == 0 : M.mv[lt + 1][h] != kKeyNone);
        if (wanted) {
#pragma unroll 4
          for (int p = part; p < n; p += parts) {
            const int cp = (staged ? (int)M.ccl[cb + p] : __ldcg(w.tc_cloud + toff + p)) & (SKYOPT_MAX_CLOUDS - 1);
            const double val = staged ? M.cval[cb + p] : __ldcg(w.tc_value + toff + p);
            const double dpp = __dadd_rn(val, M.B[lt][cp]);
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
          const int slot_t = lt + 1;  // indexed by the child task; T = the sink
          M.bk_idx[slot_t][h] = bi;
          unsigned char cl = 0;
          if (bi != 0x7FFFFFFF && bi < n)
            cl = staged ? M.ccl[cb + bi] : (unsigned char)__ldcg(w.tc_cloud + toff + bi);
          M.bk_cl[slot_t][h] = cl;
          if (sink) M.obj = bv;
        }
      }
      // back-tracking (thread 0; rehearsed by thread 32)
      if (pass == 1) { __syncthreads(); step_mark(out.trace, 4); }
      if (tid == (pass == 1 ? 0 : 32)) {
        int idx = M.bk_idx[T][0];
        int cl = M.bk_cl[T][0] & (SKYOPT_MAX_CLOUDS - 1);
#pragma unroll 1
        for (int lt = T - 1; lt >= 0; --lt) {
          if (pass == 1) M.choice[lt] = idx;
          if (lt > 0) {
            const int ni = M.bk_idx[lt][cl];
            cl = M.bk_cl[lt][cl] & (SKYOPT_MAX_CLOUDS - 1);
            idx = ni;
          }
        }
        if (pass == 1) {
          SkyoptDagResult r; r.status = 0; r.task_fail = -1; r.objective = M.obj;
          out.dag[dag] = r;
        }
      }
    }
    if (pass == 0) { __syncthreads(); step_mark(out.trace, 3); }
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
  int32_t *dag_done;      // [n_dags] tasks placed so far (zero between launches)
  unsigned int *sync;     // [2] arrivals at the barrier / at the exit (zero between launches)
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
    while (ld_acquire_u32(a.sync) < gridDim.x) if (!(a.scan.noprune & 4u)) __nanosleep(64);
    __threadfence();
  }
  __syncthreads();
  for (int t = blockIdx.x; t < a.n_tasks; t += gridDim.x) {
    place_body(a.place, t, smem_step);
    if (a.do_solve) {
      const int dag = a.task_dag[t];
      __syncthreads();  // the task's tables are written
      if (tid == 0) {
        const SkyoptDag D = a.place.in.dags[dag];
        __threadfence();
        const int done = atomicAdd(a.dag_done + dag, 1);
        s_last = (done == D.task_end - D.task_begin - 1) ? 1 : 0;
        if (s_last) a.dag_done[dag] = 0;  // ready for the next launch
      }
      __syncthreads();
      if (s_last) {
        __threadfence();
        chain_body(a.scan.cat, a.place.in, a.place.w, a.out, a.place.task_mv, dag, smem_step);
      }
    }
    __syncthreads();  // shared memory is reused by the next task
  }
  // ---- the last block out re-arms the barrier and the scan results (the
  // next launch of this context starts from "nothing found")
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(a.sync + 1, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
#pragma unroll 1
    for (int i = tid; i < a.n_queries; i += kScanThreads) { a.scan.best_rank[i] = kRankNone; a.scan.any1[i] = 0u; }
    if (a.scan.group_ready)
#pragma unroll 1
      for (int i = tid; i < a.scan.n_groups; i += kScanThreads) a.scan.group_ready[i] = 0u;
    if (tid == 0) { __threadfence(); a.sync[0] = 0; a.sync[1] = 0; }
  }
}

}  // namespace skyopt
