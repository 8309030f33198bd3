// The whole message-passing loop of a SMALL graph in one launch (sm_100a).
//
// The reference's default batch is 3 samples (code/train_options.ini:26; BASELINE config 1: RouteNet on NSFNET,
// 546 paths and 126 links per step).  At that size the T = 8 iterations of generate_model.py:405-602 are 16 dependent
// launches of kernels that each fill a fraction of the machine for a few microseconds: launch gaps and the kernels'
// own ramp-up are the whole step.  Here one persistent grid runs all T iterations: the GRU weights of every message
// passing stay in shared memory from the first iteration to the last, a warp owns a destination row (lane = unit; two
// rows per warp at 16 units), the stages are separated by a grid-wide barrier instead of a launch boundary, and the
// states ping-pong between two buffers per entity.
//
// Arithmetic: exactly the order of the fp32 kernels in gru.cuh (bias, then x K, then h R, one accumulator per gate;
// gru_out), the neighbour sum in CSR order like ign_segment_reduce -- so the result is BIT-IDENTICAL to the launch-
// per-stage fp32 path (tests/test_gpu_graphs.py), and the parity of that path carries over.
//
// Stages it runs (everything RouteNet / Q-size shaped models need, auxilary_classes.py:752-796, :254-262):
//   kind 0  ordered / interleave / concat-axis-1 update: walk the destination's step entries, one GRU step each
//   kind 1  sum aggregation over the CSR row + one GRU step
// States read across the barrier go through L2 (ld.global.cg): L1 is not coherent between SMs.

#include "gru.cuh"

namespace {

constexpr int SG_MAX_OPS = 8;
constexpr int SG_MAX_ENT = 8;
constexpr int SG_MAX_STEPS = 64;      // stages x iterations when every stage keeps its outputs (training)
constexpr int SG_THREADS = 512;       // 16 warps per SM: the stages are latency chains, more rows in flight is what pays
constexpr int SG_AHEAD = 8;           // message rows of a walk in flight ahead of the step that consumes them

struct SgOp {
  int kind;
  int dst;
  int src[IGN_MAX_SOURCES];
  const int* rowptr;
  const int* idx;
  const float* kernel;
  const float* rkernel;
  const float* bias;
};

struct SgProgram {
  int n_ops, n_ent, iterations;
  int rows[SG_MAX_ENT];
  float* buf[SG_MAX_ENT][2];
  SgOp op[SG_MAX_OPS];
  unsigned int* barrier;
  // training (tf.gradients needs every intermediate, generate_model.py:791): stage s = iteration * n_ops + o writes
  // its new states to step_out[s] instead of the entity's other buffer, the state after every step of a walk to
  // step_hseq[s] (row = position of the step entry) and the neighbour sum to step_agg[s]
  int keep;
  float* step_out[SG_MAX_STEPS];
  float* step_hseq[SG_MAX_STEPS];
  float* step_agg[SG_MAX_STEPS];
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of the grid are resident (the host sizes the grid to one CTA per SM at most)
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target) {
  target += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    // arrive without waiting for the atomic's round trip; release orders the CTA's stores (seen through bar.sync)
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    while (ld_acquire_u32(ctr) < target) {
    }
  }
  __syncthreads();
}

// One GRU step for NR rows at once (lane = unit): the weight loads are shared by the rows, which is what bounds the
// kernel once every SM holds several warps (3 shared-memory wavefronts per k for the weights, 1 per row for the shuffle)
template <int U, int NR>
__device__ __forceinline__ void gru_step(const float* __restrict__ w, const float (&x)[NR], float (&h)[NR], const bool (&act)[NR],
                                         int u) {
  const float* K = w;
  const float* R = w + U * 3 * U;
  const float* b0 = R + U * 3 * U;
  const float* b1 = b0 + 3 * U;
  float az[NR], ar[NR], axh[NR], ahh[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    az[i] = b0[u] + b1[u];
    ar[i] = b0[U + u] + b1[U + u];
    axh[i] = b0[2 * U + u];
    ahh[i] = b1[2 * U + u];
  }
#pragma unroll 8
  for (int k = 0; k < U; ++k) {
    const float wz = K[k * 3 * U + u], wr = K[k * 3 * U + U + u], wh = K[k * 3 * U + 2 * U + u];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const float xk = __shfl_sync(0xffffffffu, x[i], k, U);
      az[i] = fmaf(xk, wz, az[i]);
      ar[i] = fmaf(xk, wr, ar[i]);
      axh[i] = fmaf(xk, wh, axh[i]);
    }
  }
#pragma unroll 8
  for (int k = 0; k < U; ++k) {
    const float wz = R[k * 3 * U + u], wr = R[k * 3 * U + U + u], wh = R[k * 3 * U + 2 * U + u];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      const float hk = __shfl_sync(0xffffffffu, h[i], k, U);
      az[i] = fmaf(hk, wz, az[i]);
      ar[i] = fmaf(hk, wr, ar[i]);
      ahh[i] = fmaf(hk, wh, ahh[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < NR; ++i)
    if (act[i]) h[i] = ign_gru::gru_out(az[i], ar[i], axh[i], ahh[i], h[i]);
}

struct SgStage {
  const SgOp* op;
  const float* w;
  const float* hcur;
  float* hout;
  float* hseq;      // training: state after every step of the walk, or nullptr
  float* agg;       // training: the neighbour sum, or nullptr
  const float* tbl[IGN_MAX_SOURCES];
  int n;
};

// NR destination rows (row groups at 16 units) of one stage, walked together by one warp
template <int U, int NR>
__device__ __forceinline__ void run_rows(const SgStage& sg, const int (&first)[NR], int u, int sub, int* slot_e, int* slot_ll,
                                         bool use_cache, bool fill_cache) {
  constexpr int RPW = 32 / U;
  const SgOp& op = *sg.op;
  int lo[NR], len[NR], my_e[NR];
  float h[NR];
  bool valid[NR];
  int maxlen = 0;
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int d = first[i] + sub;
    valid[i] = d < sg.n;
    lo[i] = 0; len[i] = 0; my_e[i] = -1;
    h[i] = valid[i] ? __ldcg(sg.hcur + (int64_t)d * U + u) : 0.f;
    if (use_cache) {                               // NR == 1 only
      lo[i] = slot_ll[0]; len[i] = slot_ll[1]; my_e[i] = *slot_e;
    } else {
      if (valid[i]) {
        lo[i] = __ldg(op.rowptr + d);
        len[i] = __ldg(op.rowptr + d + 1) - lo[i];
      }
      // the row's entries U at a time, one per lane (a coalesced load), handed out by shuffles: the index load is
      // off the dependent chain of every step
      my_e[i] = u < len[i] ? __ldg(op.idx + lo[i] + u) : -1;
      if (fill_cache) {
        *slot_e = my_e[i];
        if (u == 0) { slot_ll[0] = lo[i]; slot_ll[1] = len[i]; }
      }
    }
    maxlen = max(maxlen, len[i]);
  }
  if (RPW > 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, U & 31));
  bool act[NR];
  if (op.kind == 0) {
    auto row_of = [&](int e) -> float {
      if (e < 0) return 0.f;
      const int k = e >> IGN_STEP_SRC_SHIFT;
      const float* base_p = k == 0 ? sg.tbl[0] : k == 1 ? sg.tbl[1] : k == 2 ? sg.tbl[2] : sg.tbl[3];
      return __ldcg(base_p + (int64_t)(e & IGN_STEP_ROW_MASK) * U + u);
    };
    // SG_AHEAD message rows per destination in flight: the steps then run back to back at the speed of their gates
    float xs[SG_AHEAD][NR];
#pragma unroll
    for (int q = 0; q < SG_AHEAD; ++q) {
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int e = __shfl_sync(0xffffffffu, my_e[i], q, U);
        xs[q][i] = q < len[i] ? row_of(e) : 0.f;
      }
    }
    for (int t0 = 0; t0 < maxlen; t0 += SG_AHEAD) {
#pragma unroll
      for (int q = 0; q < SG_AHEAD; ++q) {
        const int t = t0 + q;
        if (t < maxlen) {                          // uniform over the warp
          float x[NR];
          const int ta = t + SG_AHEAD;
#pragma unroll
          for (int i = 0; i < NR; ++i) {
            x[i] = xs[q][i];
            if (ta % U == 0) my_e[i] = ta + u < len[i] ? __ldg(op.idx + lo[i] + ta + u) : -1;   // next block of entries
            const int en = __shfl_sync(0xffffffffu, my_e[i], ta % U, U);
            xs[q][i] = ta < len[i] ? row_of(en) : 0.f;
            act[i] = t < len[i];
          }
          gru_step<U, NR>(sg.w, x, h, act, u);
          if (sg.hseq) {
#pragma unroll
            for (int i = 0; i < NR; ++i)
              if (act[i]) sg.hseq[(int64_t)(lo[i] + t) * U + u] = h[i];
          }
        }
      }
    }
  } else {
    float x[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      x[i] = 0.f;
      act[i] = true;
    }
    for (int t0 = 0; t0 < maxlen; t0 += U) {
      const int cnt = min(U, maxlen - t0);
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        if (t0 > 0) my_e[i] = t0 + u < len[i] ? __ldg(op.idx + lo[i] + t0 + u) : -1;
        float r[U];                                // the whole block of neighbour rows in flight, added in CSR order
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const int e = __shfl_sync(0xffffffffu, my_e[i], q, U);
          r[q] = (q < cnt && t0 + q < len[i]) ? __ldcg(sg.tbl[0] + (int64_t)e * U + u) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < U; ++q)
          if (q < cnt && t0 + q < len[i]) x[i] += r[q];
      }
    }
    if (sg.agg) {
#pragma unroll
      for (int i = 0; i < NR; ++i)
        if (valid[i]) sg.agg[(int64_t)(first[i] + sub) * U + u] = x[i];
    }
    gru_step<U, NR>(sg.w, x, h, act, u);
  }
#pragma unroll
  for (int i = 0; i < NR; ++i)
    if (valid[i]) sg.hout[(int64_t)(first[i] + sub) * U + u] = h[i];
}

template <int U>
__global__ void __launch_bounds__(SG_THREADS, 1) small_graph_kernel(const __grid_constant__ SgProgram pg) {
  constexpr int WF = 2 * U * 3 * U + 6 * U;        // K | R | bias[2][3U] per stage
  constexpr int RPW = 32 / U;                      // rows per warp
  extern __shared__ float4 sg_smem4[];
  float* sw = reinterpret_cast<float*>(sg_smem4);
  for (int o = 0; o < pg.n_ops; ++o) {
    float* w = sw + o * WF;
    for (int i = threadIdx.x * 4; i < U * 3 * U; i += SG_THREADS * 4) {
      st_f4(w + i, ldg_f4(pg.op[o].kernel + i));
      st_f4(w + U * 3 * U + i, ldg_f4(pg.op[o].rkernel + i));
    }
    for (int i = threadIdx.x; i < 6 * U; i += SG_THREADS) w[2 * U * 3 * U + i] = pg.op[o].bias[i];
  }
  __syncthreads();
  // rowptr and the first U step entries of a warp's row do not change between iterations: kept in shared memory after
  // the first one when the stage has at most one row (pair) per warp, which takes two dependent global loads off the
  // start of every later stage
  int* cache = reinterpret_cast<int*>(sw + pg.n_ops * WF);

  const int lane = threadIdx.x & 31, u = lane % U, sub = lane / U;
  // consecutive rows go to different SMs: with fewer rows than warps every SM still gets its share
  const int warp = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;
  const int n_warps = gridDim.x * (SG_THREADS / 32);
  const int stride = n_warps * RPW;
  __shared__ float* s_cur[SG_MAX_ENT];             // where every entity's current state lives
  if (threadIdx.x < SG_MAX_ENT) s_cur[threadIdx.x] = pg.buf[threadIdx.x][0];
  __syncthreads();
  unsigned target = 0;
  const int total = pg.iterations * pg.n_ops;
  for (int step = 0; step < total; ++step) {
    const int o = step % pg.n_ops;
    SgStage sg;
    sg.op = &pg.op[o];
    sg.w = sw + o * WF;
    sg.n = pg.rows[sg.op->dst];
    sg.hcur = s_cur[sg.op->dst];
    sg.hout = pg.keep ? pg.step_out[step] : (sg.hcur == pg.buf[sg.op->dst][0] ? pg.buf[sg.op->dst][1] : pg.buf[sg.op->dst][0]);
    sg.hseq = pg.keep ? pg.step_hseq[step] : nullptr;
    sg.agg = pg.keep ? pg.step_agg[step] : nullptr;
#pragma unroll
    for (int k = 0; k < IGN_MAX_SOURCES; ++k) sg.tbl[k] = sg.op->src[k] >= 0 ? s_cur[sg.op->src[k]] : nullptr;
    int* slot_e = cache + o * (SG_THREADS + SG_THREADS / 32 * 4) + threadIdx.x;
    int* slot_ll = cache + o * (SG_THREADS + SG_THREADS / 32 * 4) + SG_THREADS + ((threadIdx.x >> 5) * 2 + sub) * 2;
    if (sg.n <= stride) {                          // at most one row (pair) per warp: its metadata stays in shared memory
      const int first[1] = {warp * RPW};
      if (first[0] < sg.n) run_rows<U, 1>(sg, first, u, sub, slot_e, slot_ll, step >= pg.n_ops, step < pg.n_ops);
    } else {
      int base = warp * RPW;
      for (; base + stride < sg.n; base += 2 * stride) {       // two rows share every weight load
        const int first[2] = {base, base + stride};
        run_rows<U, 2>(sg, first, u, sub, slot_e, slot_ll, false, false);
      }
      if (base < sg.n) {
        const int first[1] = {base};
        run_rows<U, 1>(sg, first, u, sub, slot_e, slot_ll, false, false);
      }
    }
    __syncthreads();                               // every warp has read this stage's pointers
    if (threadIdx.x == 0) s_cur[sg.op->dst] = sg.hout;
#ifdef IGN_SG_PROFILE                              // CTA 0: stage start, rows done, barrier passed (ns)
    unsigned long long* prof = reinterpret_cast<unsigned long long*>(pg.barrier) + 8;
    unsigned long long t_done = 0;
    if (threadIdx.x == 0 && blockIdx.x == 0 && step < 32) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_done));
      prof[step * 2] = t_done;
    }
#endif
    if (step + 1 < total) grid_barrier(pg.barrier, target);
#ifdef IGN_SG_PROFILE
    if (threadIdx.x == 0 && blockIdx.x == 0 && step < 32) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_done));
      prof[step * 2 + 1] = t_done;
    }
#endif
  }
}

template <int U>
int launch(const SgProgram& pg, int max_rows, cudaStream_t st) {
  constexpr int WF = 2 * U * 3 * U + 6 * U;
  const size_t smem = (size_t)pg.n_ops * (WF * sizeof(float) + (SG_THREADS + SG_THREADS / 32 * 4) * sizeof(int));
  const void* fn = reinterpret_cast<const void*>(&small_graph_kernel<U>);
  static thread_local int sms[16];                 // per device: SM count once the kernel's attribute is set
  int dev = 0;
  IGN_CUDA(cudaGetDevice(&dev));
  IGN_REQUIRE(dev >= 0 && dev < 16, IGN_ERR_UNSUPPORTED, "IGNNITION: small_graph: device index above 15");
  if (!sms[dev]) {
    const int max_smem = SG_MAX_OPS * (WF * (int)sizeof(float) + (SG_THREADS + SG_THREADS / 32 * 4) * (int)sizeof(int));
    IGN_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    int n = 0, occ = 0;
    IGN_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    IGN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, small_graph_kernel<U>, SG_THREADS, max_smem));
    IGN_REQUIRE(occ >= 1, IGN_ERR_UNSUPPORTED, "IGNNITION: small_graph: the kernel does not fit an SM");
    sms[dev] = n;
  }
  const int rows_per_warp = 32 / U;                // one warp per SM first: rows are dealt out across CTAs
  int grid = (max_rows + rows_per_warp - 1) / rows_per_warp;
  grid = grid < 1 ? 1 : grid > sms[dev] ? sms[dev] : grid;      // every CTA resident: the barrier spins
  IGN_CUDA(cudaMemsetAsync(pg.barrier, 0, sizeof(unsigned int), st));
  small_graph_kernel<U><<<grid, SG_THREADS, smem, st>>>(pg);
  IGN_CHECK_LAUNCH("small_graph");
  return IGN_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// All adjacencies of a small graph -> CSR by destination in ONE launch, one CTA per adjacency (ign_csr_build is a
// dozen launches of a few microseconds each at this size).  Degrees by shared-memory atomics, one block scan, then
//   seq given : every edge at rowptr[dst] + seq (the reference's scatter_nd((dst, seq)), generate_model.py:490);
//               slots no edge claims stay -1 = a zero row for every consumer
//   no seq    : stable order of the input; edges already grouped by destination (what the generator emits,
//               generator_std_to_framework.py:134-185) are placed in parallel, anything else by one warp walking
//               the list in order (match_any groups + per-destination cursors)
constexpr int CS_THREADS = 1024;
constexpr int CS_MAX_ADJ = 8;

struct CsAdj {
  const int* dst;
  const int* src;
  const int* seq;
  int* rowptr;
  int* col;
  int* perm;
  int n_edges, num_dst;
};
struct CsList {
  CsAdj a[CS_MAX_ADJ];
};

__global__ void __launch_bounds__(CS_THREADS) csr_small_kernel(const __grid_constant__ CsList L) {
  const CsAdj& a = L.a[blockIdx.x];
  extern __shared__ int cs_smem[];
  int* cnt = cs_smem;                              // [num_dst + 1] degrees -> exclusive offsets -> cursors
  __shared__ int warp_sums[CS_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = a.n_edges, nd = a.num_dst;
  for (int i = tid; i <= nd; i += CS_THREADS) cnt[i] = 0;
  __syncthreads();
  int unsorted = 0;
  for (int e = tid; e < n; e += CS_THREADS) {
    const int d = a.dst[e];
    if (d >= 0 && d < nd) atomicAdd(&cnt[d], 1);
    if (e > 0 && a.dst[e - 1] > d) unsorted = 1;
  }
  unsorted = __syncthreads_or(unsorted);
  // exclusive scan of cnt[0 .. nd]: a contiguous run per thread, warp shuffles, one pass over the warp sums
  const int ipt = (nd + 1 + CS_THREADS - 1) / CS_THREADS;
  const int i0 = min(tid * ipt, nd + 1), i1 = min(i0 + ipt, nd + 1);
  int run = 0;
  for (int i = i0; i < i1; ++i) run += cnt[i];
  int inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    const int v = warp_sums[lane];
    int winc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_sums[lane] = winc - v;
  }
  __syncthreads();
  int off = warp_sums[wid] + inc - run;
  for (int i = i0; i < i1; ++i) {
    const int c = cnt[i];
    cnt[i] = off;
    a.rowptr[i] = off;
    off += c;
  }
  __syncthreads();
  if (a.seq) {
    for (int e = tid; e < n; e += CS_THREADS) {
      a.col[e] = -1;
      if (a.perm) a.perm[e] = -1;
    }
    __syncthreads();
    for (int e = tid; e < n; e += CS_THREADS) {
      const int d = a.dst[e];
      if (d < 0 || d >= nd) continue;
      const int lo = cnt[d], hi = cnt[d + 1];
      const int pos = lo + a.seq[e];
      if (pos < lo || pos >= hi) continue;         // malformed seq: the slot stays a zero row
      a.col[pos] = a.src[e];
      if (a.perm) a.perm[pos] = e;
    }
  } else if (!unsorted) {
    for (int e = tid; e < n; e += CS_THREADS) {
      a.col[e] = a.src[e];
      if (a.perm) a.perm[e] = e;
    }
  } else {
    // any order of the input (the transposed adjacencies of a train step).  Short rows: every edge claims a slot of
    // its row with a shared-memory atomic, then one thread per row sorts the row's edge ids -- ids are unique, so the
    // result is the stable order, bit for bit what the radix sort of ign_csr_build gives.  A row longer than 256
    // would make that sort quadratic: one warp then walks the list in order (match_any groups + cursors).
    int long_row = 0;
    for (int i = tid; i < nd; i += CS_THREADS) long_row |= (a.rowptr[i + 1] - a.rowptr[i]) > 256;
    long_row = __syncthreads_or(long_row);
    int* ids = a.perm ? a.perm : a.col;
    if (!long_row) {
      for (int e = tid; e < n; e += CS_THREADS) {
        const int d = a.dst[e];
        if (d >= 0 && d < nd) ids[atomicAdd(&cnt[d], 1)] = e;
      }
      __syncthreads();
      for (int i = tid; i < nd; i += CS_THREADS) {
        const int lo = a.rowptr[i], hi = a.rowptr[i + 1];
        for (int p = lo + 1; p < hi; ++p) {
          const int v = ids[p];
          int q = p - 1;
          while (q >= lo && ids[q] > v) {
            ids[q + 1] = ids[q];
            --q;
          }
          ids[q + 1] = v;
        }
        for (int p = lo; p < hi; ++p) a.col[p] = a.src[ids[p]];
      }
    } else if (wid == 0) {
      for (int base = 0; base < n; base += 32) {
        const int e = base + lane;
        const int d = e < n ? a.dst[e] : -1;
        const bool ok = d >= 0 && d < nd;
        const unsigned grp = __match_any_sync(0xffffffffu, ok ? d : -1 - lane);
        const int leader = __ffs(grp) - 1;
        int start = 0;
        if (ok && lane == leader) {
          start = cnt[d];
          cnt[d] = start + __popc(grp);
        }
        start = __shfl_sync(0xffffffffu, start, leader);
        if (ok) {
          const int pos = start + __popc(grp & ((1u << lane) - 1u));
          a.col[pos] = a.src[e];
          if (a.perm) a.perm[pos] = e;
        }
        __syncwarp();
      }
    }
  }
}

}  // namespace

extern "C" size_t ign_small_graph_ws_bytes(void) { return 1024; }    // barrier counter (+ stage timestamps when profiling)

extern "C" int ign_small_graph_forward(int units, int n_entities, const int64_t* rows, float* const* buf0,
                                       float* const* buf1, int n_ops, const int32_t* op_kind, const int32_t* op_dst,
                                       const int32_t* op_src, const int32_t* const* op_rowptr,
                                       const int32_t* const* op_idx, const float* const* op_kernel,
                                       const float* const* op_rkernel, const float* const* op_bias, int iterations,
                                       float* const* step_out, float* const* step_hseq, float* const* step_agg,
                                       int32_t* final_buffer, void* ws, size_t ws_bytes, void* stream) {
  IGN_REQUIRE(units == 16 || units == 32, IGN_ERR_UNSUPPORTED, "IGNNITION: small_graph: 16 or 32 units, got %d", units);
  IGN_REQUIRE(n_entities >= 1 && n_entities <= SG_MAX_ENT && n_ops >= 1 && n_ops <= SG_MAX_OPS && iterations >= 0,
              IGN_ERR_UNSUPPORTED, "IGNNITION: small_graph: at most %d entities and %d stages", SG_MAX_ENT, SG_MAX_OPS);
  IGN_REQUIRE(rows && buf0 && buf1 && op_kind && op_dst && op_src && op_rowptr && op_idx && op_kernel && op_rkernel &&
                  op_bias && final_buffer,
              IGN_ERR_INVALID, "IGNNITION: small_graph: null pointer");
  IGN_REQUIRE(ws && ws_bytes >= sizeof(unsigned int), IGN_ERR_WORKSPACE, "IGNNITION: small_graph: workspace too small");
  IGN_REQUIRE(!step_out || ((int64_t)iterations * n_ops <= SG_MAX_STEPS && step_hseq && step_agg), IGN_ERR_UNSUPPORTED,
              "IGNNITION: small_graph: keeping every stage's outputs is built for at most %d stages x iterations",
              SG_MAX_STEPS);
  SgProgram pg;
  pg.keep = step_out ? 1 : 0;
  for (int i = 0; i < SG_MAX_STEPS; ++i) {
    const bool used = step_out && i < iterations * n_ops;
    pg.step_out[i] = used ? step_out[i] : nullptr;
    pg.step_hseq[i] = used ? step_hseq[i] : nullptr;
    pg.step_agg[i] = used ? step_agg[i] : nullptr;
  }
  pg.n_ops = n_ops;
  pg.n_ent = n_entities;
  pg.iterations = iterations;
  pg.barrier = static_cast<unsigned int*>(ws);
  int64_t max_rows = 0;
  for (int e = 0; e < SG_MAX_ENT; ++e) {
    pg.rows[e] = 0;
    pg.buf[e][0] = pg.buf[e][1] = nullptr;
  }
  for (int e = 0; e < n_entities; ++e) {
    IGN_REQUIRE(rows[e] >= 0 && rows[e] <= IGN_STEP_ROW_MASK, IGN_ERR_INVALID, "IGNNITION: small_graph: bad row count");
    IGN_REQUIRE(rows[e] == 0 || (buf0[e] && buf1[e]), IGN_ERR_INVALID, "IGNNITION: small_graph: null state buffer");
    pg.rows[e] = (int)rows[e];
    pg.buf[e][0] = buf0[e];
    pg.buf[e][1] = buf1[e];
    final_buffer[e] = 0;
  }
  for (int o = 0; o < n_ops; ++o) {
    SgOp& op = pg.op[o];
    op.kind = op_kind[o];
    op.dst = op_dst[o];
    IGN_REQUIRE((op.kind == 0 || op.kind == 1) && op.dst >= 0 && op.dst < n_entities, IGN_ERR_INVALID,
                "IGNNITION: small_graph: bad stage %d", o);
    for (int k = 0; k < IGN_MAX_SOURCES; ++k) {
      op.src[k] = op_src[o * IGN_MAX_SOURCES + k];
      IGN_REQUIRE(op.src[k] < n_entities, IGN_ERR_INVALID, "IGNNITION: small_graph: bad source entity");
    }
    IGN_REQUIRE(op.src[0] >= 0, IGN_ERR_INVALID, "IGNNITION: small_graph: a stage needs a source");
    IGN_REQUIRE(op_kernel[o] && op_rkernel[o] && op_bias[o], IGN_ERR_INVALID, "IGNNITION: small_graph: null weights");
    IGN_REQUIRE(pg.rows[op.dst] == 0 || op_rowptr[o], IGN_ERR_INVALID, "IGNNITION: small_graph: null row pointer");
    op.rowptr = op_rowptr[o];
    op.idx = op_idx[o];
    op.kernel = op_kernel[o];
    op.rkernel = op_rkernel[o];
    op.bias = op_bias[o];
    if (pg.rows[op.dst] > max_rows) max_rows = pg.rows[op.dst];
    final_buffer[op.dst] ^= iterations & 1;
    for (int it = 0; step_out && it < iterations; ++it)
      IGN_REQUIRE(pg.rows[op.dst] == 0 || step_out[it * n_ops + o], IGN_ERR_INVALID,
                  "IGNNITION: small_graph: null output buffer of stage %d", it * n_ops + o);
  }
  if (iterations == 0 || max_rows == 0) {
    for (int e = 0; e < n_entities; ++e) final_buffer[e] = 0;
    return IGN_OK;
  }
  cudaStream_t st = ign_stream(stream);
  return units == 32 ? launch<32>(pg, (int)max_rows, st) : launch<16>(pg, (int)max_rows, st);
}

extern "C" int ign_csr_build_small(int n_adj, const int32_t* const* dst, const int32_t* const* src,
                                   const int32_t* const* seq, const int64_t* n_edges, const int64_t* num_dst,
                                   int32_t* const* rowptr, int32_t* const* col, int32_t* const* perm, void* stream) {
  IGN_REQUIRE(n_adj >= 1 && n_adj <= CS_MAX_ADJ, IGN_ERR_UNSUPPORTED, "IGNNITION: csr_build_small: 1..%d adjacencies", CS_MAX_ADJ);
  IGN_REQUIRE(dst && src && seq && n_edges && num_dst && rowptr && col && perm, IGN_ERR_INVALID,
              "IGNNITION: csr_build_small: null pointer");
  CsList L;
  int64_t max_dst = 0;
  for (int k = 0; k < n_adj; ++k) {
    IGN_REQUIRE(n_edges[k] >= 0 && n_edges[k] < (int64_t)1 << 30 && num_dst[k] >= 0 && num_dst[k] <= 11000, IGN_ERR_UNSUPPORTED,
                "IGNNITION: csr_build_small: at most 11000 destinations per adjacency (use ign_csr_build)");
    IGN_REQUIRE(rowptr[k] && (n_edges[k] == 0 || (dst[k] && src[k] && col[k])), IGN_ERR_INVALID,
                "IGNNITION: csr_build_small: null array");
    L.a[k] = CsAdj{dst[k], src[k], seq[k], rowptr[k], col[k], perm[k], (int)n_edges[k], (int)num_dst[k]};
    if (num_dst[k] > max_dst) max_dst = num_dst[k];
  }
  csr_small_kernel<<<n_adj, CS_THREADS, (size_t)(max_dst + 1) * sizeof(int), ign_stream(stream)>>>(L);
  IGN_CHECK_LAUNCH("csr_small");
  return IGN_OK;
}
