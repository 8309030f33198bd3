// Adjacency -> CSR by destination on the device (sm_100a).
//
// Replaces the reference's host per-edge loop and dense padded scatter
// (code/utils/generator_std_to_framework.py:134-185, code/utils/generate_model.py:479-490):
//   * stable LSD radix sort of the edge list by destination (8-bit digits, match-based warp
//     ranking, one histogram + scan + scatter per digit),
//   * rowptr from the sorted keys by boundary fill (no atomics, deterministic),
//   * or, when `seq` is given, histogram + scan + placement at rowptr[dst] + seq,
//   * destinations ordered by length for the sequence kernel, and the step table of the
//     multi-source ordered / interleave aggregation.
// All integer work: HBM-bound, no tensor cores.  Bytes per edge are stated in DESIGN.md.

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// device-wide exclusive scan (int32), multi-level
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = winc - w;   // exclusive warp offsets
    if (lane == SCAN_THREADS / 32 - 1) *total = winc;
  }
  __syncthreads();
  return warp_sums[warp] + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const int* __restrict__ in, int64_t n,
                                                                   int* __restrict__ sums) {
  __shared__ int total;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) s += in[base + i];
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// out[i] = offsets[block] + exclusive prefix inside the tile.  in == out is allowed.
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const int* in, int* out, int64_t n,
                                                                  const int* __restrict__ offsets) {
  __shared__ int total;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  int run = block_exclusive_scan(s, &total) + (offsets ? offsets[blockIdx.x] : 0);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

size_t scan_ws_bytes(int64_t n) {
  size_t total = 0;
  while (n > SCAN_TILE) {
    n = ign_cdiv(n, SCAN_TILE);
    total += ign_align((size_t)n * sizeof(int));
  }
  return total;
}

int exclusive_scan(const int* in, int* out, int64_t n, void* ws, cudaStream_t st) {
  if (n <= 0) return IGN_OK;
  if (n <= SCAN_TILE) {
    scan_apply_kernel<<<1, SCAN_THREADS, 0, st>>>(in, out, n, nullptr);
    IGN_CHECK_LAUNCH("scan_apply");
    return IGN_OK;
  }
  const int64_t nb = ign_cdiv(n, SCAN_TILE);
  int* sums = reinterpret_cast<int*>(ws);
  void* next = reinterpret_cast<char*>(ws) + ign_align((size_t)nb * sizeof(int));
  scan_reduce_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, sums);
  IGN_CHECK_LAUNCH("scan_reduce");
  int rc = exclusive_scan(sums, sums, nb, next, st);
  if (rc) return rc;
  scan_apply_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n, sums);
  IGN_CHECK_LAUNCH("scan_apply");
  return IGN_OK;
}

// ------------------------------------------------------------------------------------------
// LSD radix sort, 8-bit digits, key = uint32, value = int32 (edge id)
// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 4096 keys per CTA
constexpr int RADIX = 256;

// per-CTA digit histogram, written bin-major (hist[d * nblocks + b]) so one scan yields the
// global base of every (digit, CTA) pair
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint32_t* __restrict__ keys, int64_t n,
                                                             int shift, int* __restrict__ hist, int nblocks,
                                                             const int* __restrict__ unsorted) {
  __shared__ int h[RADIX];
  if (unsorted && *unsorted == 0) return;              // input already in key order: nothing to sort
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int i = 0; i < RS_ITEMS; ++i) {
    int64_t idx = base + i * RS_THREADS + threadIdx.x;
    if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & (RADIX - 1)], 1);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// stable scatter: warp w of the CTA owns keys [tile + w*512, tile + (w+1)*512); item i of lane l is
// key chunk + i*32 + l, so (item, lane) order == input order.  __match_any_sync groups equal
// digits; the lowest lane of each group bumps the warp's running counter.
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint32_t* __restrict__ keys_in,
                                                                const int* __restrict__ vals_in,
                                                                uint32_t* __restrict__ keys_out,
                                                                int* __restrict__ vals_out, int64_t n, int shift,
                                                                const int* __restrict__ offsets, int nblocks,
                                                                const int* __restrict__ unsorted) {
  __shared__ int cnt[RS_WARPS][RADIX];
  if (unsorted && *unsorted == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_WARPS * RADIX; i += RS_THREADS) (&cnt[0][0])[i] = 0;
  __syncthreads();

  const int64_t chunk = (int64_t)blockIdx.x * RS_TILE + warp * (32 * RS_ITEMS);
  uint32_t key[RS_ITEMS];
  int rank[RS_ITEMS];
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const int64_t idx = chunk + i * 32 + lane;
    const bool valid = idx < n;
    key[i] = valid ? keys_in[idx] : 0xffffffffu;
    const int digit = valid ? (int)((key[i] >> shift) & (RADIX - 1)) : RADIX;   // RADIX = "no key"
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const int leader = __ffs(peers) - 1;
    int old = 0;
    if (valid && lane == leader) {
      old = cnt[warp][digit];
      cnt[warp][digit] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[i] = old + __popc(peers & lt_mask);
    __syncwarp();
  }
  __syncthreads();
  {   // thread d turns the per-warp counts of digit d into global bases
    const int d = threadIdx.x;
    int run = offsets[(int64_t)d * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      int c = cnt[w][d];
      cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const int64_t idx = chunk + i * 32 + lane;
    if (idx < n) {
      const int digit = (int)((key[i] >> shift) & (RADIX - 1));
      const int pos = cnt[warp][digit] + rank[i];
      keys_out[pos] = key[i];
      vals_out[pos] = vals_in ? vals_in[idx] : (int)idx;
    }
  }
}

struct SortPlan {
  int passes;
  int nblocks;
  size_t keys_bytes, hist_bytes, scan_bytes;
};

SortPlan sort_plan(int64_t n, int64_t max_key_exclusive) {
  SortPlan p;
  int bits = 1;
  while (bits < 32 && ((int64_t)1 << bits) < max_key_exclusive) ++bits;
  p.passes = (bits + 7) / 8;
  p.nblocks = (int)ign_cdiv(n > 0 ? n : 1, RS_TILE);
  p.keys_bytes = ign_align((size_t)(n > 0 ? n : 1) * 4);
  p.hist_bytes = ign_align((size_t)RADIX * p.nblocks * 4);
  p.scan_bytes = scan_ws_bytes((int64_t)RADIX * p.nblocks);
  return p;
}
size_t sort_ws_bytes(int64_t n, int64_t max_key_exclusive) {
  SortPlan p = sort_plan(n, max_key_exclusive);
  return 4 * p.keys_bytes + p.hist_bytes + p.scan_bytes;   // keys x2, vals x2
}

// sorts (keys, iota) by key; the sorted keys/values end in *keys_sorted / *vals_sorted (inside ws)
// unsorted (optional, device): 0 = the keys are already non-decreasing, the passes return at once and the
// caller uses (keys, iota) instead of the buffers
int radix_sort_pairs(const uint32_t* keys, int64_t n, int64_t max_key_exclusive, void* ws,
                     const uint32_t** keys_sorted, const int** vals_sorted, cudaStream_t st,
                     const int* unsorted = nullptr) {
  SortPlan p = sort_plan(n, max_key_exclusive);
  char* w = reinterpret_cast<char*>(ws);
  uint32_t* kbuf[2] = {reinterpret_cast<uint32_t*>(w), reinterpret_cast<uint32_t*>(w + p.keys_bytes)};
  int* vbuf[2] = {reinterpret_cast<int*>(w + 2 * p.keys_bytes), reinterpret_cast<int*>(w + 3 * p.keys_bytes)};
  int* hist = reinterpret_cast<int*>(w + 4 * p.keys_bytes);
  void* scan_ws = w + 4 * p.keys_bytes + p.hist_bytes;
  const uint32_t* kin = keys;
  const int* vin = nullptr;   // iota on the first pass
  for (int pass = 0; pass < p.passes; ++pass) {
    const int shift = pass * 8;
    rs_hist_kernel<<<p.nblocks, RS_THREADS, 0, st>>>(kin, n, shift, hist, p.nblocks, unsorted);
    IGN_CHECK_LAUNCH("rs_hist");
    int rc = exclusive_scan(hist, hist, (int64_t)RADIX * p.nblocks, scan_ws, st);
    if (rc) return rc;
    rs_scatter_kernel<<<p.nblocks, RS_THREADS, 0, st>>>(kin, vin, kbuf[pass & 1], vbuf[pass & 1], n, shift, hist,
                                                         p.nblocks, unsorted);
    IGN_CHECK_LAUNCH("rs_scatter");
    kin = kbuf[pass & 1];
    vin = vbuf[pass & 1];
  }
  *keys_sorted = kin;
  *vals_sorted = vin;
  return IGN_OK;
}

// ------------------------------------------------------------------------------------------
// CSR assembly
// ------------------------------------------------------------------------------------------
// *unsorted = 1 if some key is smaller than its predecessor (flag cleared by the caller).  The generator
// emits every adjacency destination by destination (generator_std_to_framework.py:140-160), so edge lists
// usually arrive in destination order and the stable sort is the identity.
__global__ void sorted_check_kernel(const uint32_t* __restrict__ keys, int64_t n, int* __restrict__ unsorted) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 < n && keys[i] > keys[i + 1]) *unsorted = 1;
}

// rowptr[d] = number of sorted keys < d: position i owns every d in (keys[i-1], keys[i]]
// (keys = keys_orig when the input was already sorted)
__global__ void rowptr_fill_kernel(const uint32_t* __restrict__ keys_sorted, const uint32_t* __restrict__ keys_orig,
                                   const int* __restrict__ unsorted, int64_t n, int64_t num_dst,
                                   int* __restrict__ rowptr) {
  const uint32_t* __restrict__ keys = (unsorted && *unsorted == 0) ? keys_orig : keys_sorted;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const int64_t lo = (i == 0) ? 0 : (int64_t)keys[i - 1] + 1;
  int64_t hi = (i == n) ? num_dst : (int64_t)keys[i];
  if (hi > num_dst) hi = num_dst;          // out-of-range destination: caught by the status check
  for (int64_t d = lo; d <= hi; ++d) rowptr[d] = (int)i;
}

__global__ void gather_col_kernel(const int* __restrict__ perm, const int* __restrict__ src, int64_t n,
                                  int* __restrict__ col, int* __restrict__ perm_out,
                                  const int* __restrict__ unsorted) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int e = (unsorted && *unsorted == 0) ? (int)i : perm[i];
  col[i] = src[e];
  if (perm_out) perm_out[i] = e;
}

__global__ void degree_hist_kernel(const int* __restrict__ dst, int64_t n, int64_t num_dst,
                                   int* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int d = dst[i];
    if (d >= 0 && d < num_dst) atomicAdd(&counts[d], 1);
  }
}

__global__ void rank_place_kernel(const int* __restrict__ dst, const int* __restrict__ src,
                                  const int* __restrict__ seq, int64_t n, int64_t num_dst,
                                  const int* __restrict__ rowptr, int* __restrict__ col,
                                  int* __restrict__ perm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = dst[i];
  if (d < 0 || d >= num_dst) return;
  const int lo = rowptr[d], hi = rowptr[d + 1];
  int pos = lo + seq[i];
  if (pos < lo || pos >= hi) return;      // malformed seq: reported by the status check
  col[pos] = src[i];
  perm[pos] = (int)i;
}

// status[0] += slots whose seq != slot - rowptr[dst]; status[1] = max in-degree
__global__ void csr_check_kernel(const int* __restrict__ perm, const int* __restrict__ dst,
                                 const int* __restrict__ seq, const int* __restrict__ rowptr, int64_t n,
                                 int64_t num_dst, int* __restrict__ status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0, deg = 0;
  if (i < n) {
    const int e = perm[i];
    if (e < 0 || e >= n) {
      bad = 1;
    } else {
      const int d = dst[e];
      if (d < 0 || d >= num_dst) bad = 1;
      else if (seq) bad = (seq[e] != (int)(i - rowptr[d])) ? 1 : 0;
    }
  }
  if (i < num_dst) deg = rowptr[i + 1] - rowptr[i];
  bad = __reduce_add_sync(0xffffffffu, bad);
  deg = __reduce_max_sync(0xffffffffu, deg);
  if ((threadIdx.x & 31) == 0) {
    if (bad) atomicAdd(&status[0], bad);
    if (deg) atomicMax(&status[1], deg);
  }
}

__global__ void fill_int_kernel(int* p, int64_t n, int v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void length_key_kernel(const int* __restrict__ rowptr, int64_t n, uint32_t* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int len = rowptr[i + 1] - rowptr[i];
  keys[i] = 65535u - (uint32_t)min(len, 65535);   // descending length, stable
}

__global__ void copy_int_kernel(const int* __restrict__ in, int* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

struct StepSources {
  const int* rowptr[IGN_MAX_SOURCES];
  const int* col[IGN_MAX_SOURCES];
  int n;
};

__global__ void steps_len_kernel(StepSources s, int64_t num_dst, int* __restrict__ lens) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d > num_dst) return;
  int len = 0;
  if (d < num_dst)
    for (int k = 0; k < s.n; ++k) len += s.rowptr[k][d + 1] - s.rowptr[k][d];
  lens[d] = len;
}

__global__ void steps_fill_kernel(StepSources s, const int* __restrict__ dst_sample,
                                  const int* __restrict__ pos_off, const int* __restrict__ pos_src,
                                  const int* __restrict__ pos_col, int64_t num_dst,
                                  const int* __restrict__ steps_rowptr, int* __restrict__ steps) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= num_dst) return;
  const int smp = dst_sample ? dst_sample[d] : 0;
  const int p0 = pos_off[smp], p1 = pos_off[smp + 1];
  const int lo = steps_rowptr[d], hi = steps_rowptr[d + 1];
  for (int t = 0; t < hi - lo; ++t) {
    int entry = IGN_STEP_ZERO;
    if (p0 + t < p1) {
      const int k = pos_src[p0 + t], c = pos_col[p0 + t];
      const int r0 = s.rowptr[k][d], r1 = s.rowptr[k][d + 1];
      if (c < r1 - r0) entry = (k << IGN_STEP_SRC_SHIFT) | (s.col[k][r0 + c] & IGN_STEP_ROW_MASK);
    }
    steps[lo + t] = entry;
  }
}

// key of step i for the transposed (by source row) CSR of source `src_id`; foreign / zero steps go to
// the dummy bucket n_rows
__global__ void steps_keys_kernel(const int* __restrict__ steps, int64_t n, int src_id, int n_rows,
                                  int* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int e = steps[i];
  keys[i] = (e >= 0 && (e >> IGN_STEP_SRC_SHIFT) == src_id) ? (e & IGN_STEP_ROW_MASK) : n_rows;
}

// per sorted position i: (destination, first step, number of steps, first step entry)
__global__ void seq_meta_kernel(const int* __restrict__ steps_rowptr, const int* __restrict__ steps,
                                const int* __restrict__ order, int64_t n, int4* __restrict__ meta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = order ? order[i] : (int)i;
  const int lo = steps_rowptr[d], len = steps_rowptr[d + 1] - lo;
  meta[i] = make_int4(d, lo, len, len > 0 ? steps[lo] : IGN_STEP_ZERO);
}

// ---- step-major plan of an ordered aggregation (destinations sorted by descending length) ----
// cnt[l] = number of destinations with exactly l steps (l clamped to max_steps)
__global__ void seq_len_hist_kernel(const int4* __restrict__ meta, int64_t n, int max_steps, int* __restrict__ cnt) {
  __shared__ int h[1025];                              // max_steps <= 1024: per-CTA histogram first
  for (int k = threadIdx.x; k <= max_steps; k += blockDim.x) h[k] = 0;
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&h[min(meta[i].z, max_steps)], 1);
  __syncthreads();
  for (int k = threadIdx.x; k <= max_steps; k += blockDim.x)
    if (h[k]) atomicAdd(&cnt[k], h[k]);
}
// nt[t] = #destinations with more than t steps (a prefix of the sorted order); off[t] = sum_{u<t} nt[u]
__global__ void seq_plan_scan_kernel(const int* __restrict__ cnt, int max_steps, int* __restrict__ nt,
                                     int* __restrict__ off) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int run = 0;
  for (int t = max_steps - 1; t >= 0; --t) {        // suffix sums
    run += cnt[t + 1];
    nt[t] = run;
  }
  int o = 0;
  for (int t = 0; t < max_steps; ++t) { off[t] = o; o += nt[t]; }
  off[max_steps] = o;
}
// steps_T[off[t] + i] = step t of the i-th destination of the sorted order
__global__ void seq_steps_transpose_kernel(const int4* __restrict__ meta, const int* __restrict__ steps, int64_t n,
                                           int max_steps, const int* __restrict__ off, int* __restrict__ steps_T) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 m = meta[i];
  const int len = min(m.z, max_steps);
  for (int t = 0; t < len; ++t) steps_T[off[t] + i] = steps[m.y + t];
}

inline unsigned grid1d(int64_t n, int threads = 256) { return (unsigned)ign_cdiv(n > 0 ? n : 1, threads); }

}  // namespace

__global__ void range_check_kernel(const int* __restrict__ idx, int64_t n, int64_t bound, int* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int b = (i < n && (idx[i] < 0 || idx[i] >= bound)) ? 1 : 0;
  b = __reduce_add_sync(0xffffffffu, b);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(bad, b);
}

// out[pos[i]] = i for every flagged i; *count = number of flags
__global__ void compact_kernel(const int* __restrict__ flags, const int* __restrict__ pos, int64_t n, int add,
                               int* __restrict__ out, int* __restrict__ count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flags[i]) out[pos[i]] = (int)i + add;
  if (i == n - 1) *count = pos[i] + (flags[i] ? 1 : 0);
}

// ------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------
// *bad += number of idx[i] outside [0, bound): the source rows of a user-supplied adjacency (the gathers
// trust them), checked on request like the status pass of ign_csr_build
extern "C" int ign_index_range_check(const int32_t* idx, int64_t n, int64_t bound, int32_t* bad, void* stream) {
  IGN_REQUIRE(n >= 0 && bad, IGN_ERR_INVALID, "IGNNITION: index_range_check: bad argument");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(idx, IGN_ERR_INVALID, "IGNNITION: index_range_check: null pointer");
  range_check_kernel<<<grid1d(n), 256, 0, ign_stream(stream)>>>(idx, n, bound, bad);
  IGN_CHECK_LAUNCH("range_check");
  return IGN_OK;
}

extern "C" size_t ign_flag_compact_ws_bytes(int64_t n) {
  if (n < 0) return 0;
  return ign_align((size_t)(n > 0 ? n : 1) * 4) + scan_ws_bytes(n) + 256;
}

extern "C" int ign_flag_compact(const int32_t* flags, int64_t n, int add, int32_t* out, int32_t* count, void* ws,
                                size_t ws_bytes, void* stream) {
  IGN_REQUIRE(n >= 0 && count, IGN_ERR_INVALID, "IGNNITION: flag_compact: bad argument");
  cudaStream_t st = ign_stream(stream);
  if (n == 0) {
    IGN_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st));
    return IGN_OK;
  }
  IGN_REQUIRE(flags && out, IGN_ERR_INVALID, "IGNNITION: flag_compact: null pointer");
  IGN_REQUIRE(ws && ws_bytes >= ign_flag_compact_ws_bytes(n), IGN_ERR_WORKSPACE,
              "IGNNITION: flag_compact: workspace too small");
  int* pos = reinterpret_cast<int*>(ws);
  void* scan_ws = reinterpret_cast<char*>(ws) + ign_align((size_t)n * 4);
  int rc = exclusive_scan(flags, pos, n, scan_ws, st);
  if (rc) return rc;
  compact_kernel<<<grid1d(n), 256, 0, st>>>(flags, pos, n, add, out, count);
  IGN_CHECK_LAUNCH("compact");
  return IGN_OK;
}

extern "C" size_t ign_csr_build_ws_bytes(int64_t n_edges, int64_t num_dst) {
  if (n_edges < 0 || num_dst < 0) return 0;
  size_t sort = sort_ws_bytes(n_edges, num_dst > 0 ? num_dst : 1);
  size_t rank = ign_align((size_t)(num_dst + 1) * 4) + scan_ws_bytes(num_dst + 1);
  size_t perm = ign_align((size_t)(n_edges > 0 ? n_edges : 1) * 4);
  return (sort > rank ? sort : rank) + perm + 256;
}

extern "C" int ign_csr_build(const int32_t* dst, const int32_t* src, const int32_t* seq, int64_t n_edges,
                             int64_t num_dst, int mode, int32_t* rowptr, int32_t* col, int32_t* perm,
                             int32_t* status, void* ws, size_t ws_bytes, void* stream) {
  IGN_REQUIRE(n_edges >= 0 && num_dst >= 0, IGN_ERR_INVALID, "IGNNITION: csr_build: negative size");
  IGN_REQUIRE(n_edges < (int64_t)1 << 31 && num_dst < ((int64_t)1 << 31) - 1, IGN_ERR_UNSUPPORTED,
              "IGNNITION: csr_build: int32 indices only (n_edges, num_dst < 2^31)");
  IGN_REQUIRE(rowptr && (n_edges == 0 || (dst && src && col)), IGN_ERR_INVALID,
              "IGNNITION: csr_build: null pointer");
  IGN_REQUIRE(mode == IGN_CSR_SORT || mode == IGN_CSR_RANK, IGN_ERR_INVALID, "IGNNITION: csr_build: bad mode");
  IGN_REQUIRE(mode == IGN_CSR_SORT || seq || n_edges == 0, IGN_ERR_INVALID,
              "IGNNITION: csr_build: IGN_CSR_RANK needs seq");
  IGN_REQUIRE(ws && ws_bytes >= ign_csr_build_ws_bytes(n_edges, num_dst), IGN_ERR_WORKSPACE,
              "IGNNITION: csr_build: workspace too small (%zu < %zu)", ws_bytes,
              ign_csr_build_ws_bytes(n_edges, num_dst));
  cudaStream_t st = ign_stream(stream);
  if (status) IGN_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int), st));
  if (n_edges == 0) {
    fill_int_kernel<<<grid1d(num_dst + 1), 256, 0, st>>>(rowptr, num_dst + 1, 0);
    IGN_CHECK_LAUNCH("fill_int");
    return IGN_OK;
  }
  char* w = reinterpret_cast<char*>(ws);
  const size_t perm_bytes = ign_align((size_t)n_edges * 4);
  int* perm_buf = perm ? perm : reinterpret_cast<int*>(w);
  char* w2 = w + perm_bytes;

  if (mode == IGN_CSR_SORT) {
    const uint32_t* ks = nullptr;
    const int* vs = nullptr;
    // the 256 spare bytes at the end of the workspace hold the "unsorted" flag
    int* unsorted = reinterpret_cast<int*>(w + ign_csr_build_ws_bytes(n_edges, num_dst) - 256);
    const uint32_t* keys = reinterpret_cast<const uint32_t*>(dst);
    IGN_CUDA(cudaMemsetAsync(unsorted, 0, sizeof(int), st));
    sorted_check_kernel<<<grid1d(n_edges), 256, 0, st>>>(keys, n_edges, unsorted);
    IGN_CHECK_LAUNCH("sorted_check");
    int rc = radix_sort_pairs(keys, n_edges, num_dst, w2, &ks, &vs, st, unsorted);
    if (rc) return rc;
    rowptr_fill_kernel<<<grid1d(n_edges + 1), 256, 0, st>>>(ks, keys, unsorted, n_edges, num_dst, rowptr);
    IGN_CHECK_LAUNCH("rowptr_fill");
    gather_col_kernel<<<grid1d(n_edges), 256, 0, st>>>(vs, src, n_edges, col, perm_buf, unsorted);
    IGN_CHECK_LAUNCH("gather_col");
  } else {
    int* counts = reinterpret_cast<int*>(w2);
    void* scan_ws = w2 + ign_align((size_t)(num_dst + 1) * 4);
    IGN_CUDA(cudaMemsetAsync(counts, 0, (size_t)(num_dst + 1) * 4, st));
    degree_hist_kernel<<<grid1d(n_edges), 256, 0, st>>>(dst, n_edges, num_dst, counts);
    IGN_CHECK_LAUNCH("degree_hist");
    int rc = exclusive_scan(counts, rowptr, num_dst + 1, scan_ws, st);
    if (rc) return rc;
    IGN_CUDA(cudaMemsetAsync(perm_buf, 0xff, (size_t)n_edges * 4, st));   // -1 = unfilled slot
    // a slot no edge claims (gaps / duplicates in seq) is a zero row for every consumer, as in the reference's
    // scatter_nd (generate_model.py:490), never an uninitialised index
    IGN_CUDA(cudaMemsetAsync(col, 0xff, (size_t)n_edges * 4, st));
    rank_place_kernel<<<grid1d(n_edges), 256, 0, st>>>(dst, src, seq, n_edges, num_dst, rowptr, col, perm_buf);
    IGN_CHECK_LAUNCH("rank_place");
  }
  if (status) {
    const int64_t m = n_edges > num_dst ? n_edges : num_dst;
    csr_check_kernel<<<grid1d(m), 256, 0, st>>>(perm_buf, dst, seq, rowptr, n_edges, num_dst, status);
    IGN_CHECK_LAUNCH("csr_check");
  }
  return IGN_OK;
}

extern "C" size_t ign_length_order_ws_bytes(int64_t num_dst) {
  if (num_dst < 0) return 0;
  return ign_align((size_t)(num_dst > 0 ? num_dst : 1) * 4) + sort_ws_bytes(num_dst, 65536) + 256;
}

extern "C" int ign_length_order(const int32_t* rowptr, int64_t num_dst, int32_t* order, void* ws,
                                size_t ws_bytes, void* stream) {
  IGN_REQUIRE(num_dst >= 0 && num_dst < (int64_t)1 << 31, IGN_ERR_INVALID, "IGNNITION: length_order: bad size");
  if (num_dst == 0) return IGN_OK;
  IGN_REQUIRE(rowptr && order, IGN_ERR_INVALID, "IGNNITION: length_order: null pointer");
  IGN_REQUIRE(ws && ws_bytes >= ign_length_order_ws_bytes(num_dst), IGN_ERR_WORKSPACE,
              "IGNNITION: length_order: workspace too small");
  cudaStream_t st = ign_stream(stream);
  char* w = reinterpret_cast<char*>(ws);
  uint32_t* keys = reinterpret_cast<uint32_t*>(w);
  char* w2 = w + ign_align((size_t)num_dst * 4);
  length_key_kernel<<<grid1d(num_dst), 256, 0, st>>>(rowptr, num_dst, keys);
  IGN_CHECK_LAUNCH("length_key");
  const uint32_t* ks = nullptr;
  const int* vs = nullptr;
  int rc = radix_sort_pairs(keys, num_dst, 65536, w2, &ks, &vs, st);
  if (rc) return rc;
  copy_int_kernel<<<grid1d(num_dst), 256, 0, st>>>(vs, order, num_dst);
  IGN_CHECK_LAUNCH("copy_int");
  return IGN_OK;
}

extern "C" size_t ign_steps_build_ws_bytes(int64_t num_dst) {
  if (num_dst < 0) return 0;
  return ign_align((size_t)(num_dst + 1) * 4) + scan_ws_bytes(num_dst + 1) + 256;
}

extern "C" int ign_steps_build(int n_src, const int32_t* const* rowptrs, const int32_t* const* cols,
                               const int32_t* dst_sample, const int32_t* pos_off, const int32_t* pos_src,
                               const int32_t* pos_col, int64_t num_dst, int32_t* steps_rowptr, int32_t* steps,
                               void* ws, size_t ws_bytes, void* stream) {
  IGN_REQUIRE(n_src >= 1 && n_src <= IGN_MAX_SOURCES, IGN_ERR_UNSUPPORTED,
              "IGNNITION: steps_build: between 1 and %d sources are supported", IGN_MAX_SOURCES);
  IGN_REQUIRE(num_dst >= 0 && rowptrs && cols && steps_rowptr && pos_off && pos_src && pos_col, IGN_ERR_INVALID,
              "IGNNITION: steps_build: null pointer");
  IGN_REQUIRE(ws && ws_bytes >= ign_steps_build_ws_bytes(num_dst), IGN_ERR_WORKSPACE,
              "IGNNITION: steps_build: workspace too small");
  cudaStream_t st = ign_stream(stream);
  StepSources s;
  s.n = n_src;
  for (int k = 0; k < IGN_MAX_SOURCES; ++k) {
    s.rowptr[k] = k < n_src ? rowptrs[k] : nullptr;
    s.col[k] = k < n_src ? cols[k] : nullptr;
    IGN_REQUIRE(k >= n_src || (s.rowptr[k] && (s.col[k] || true)), IGN_ERR_INVALID,
                "IGNNITION: steps_build: null source CSR");
  }
  char* w = reinterpret_cast<char*>(ws);
  int* lens = reinterpret_cast<int*>(w);
  void* scan_ws = w + ign_align((size_t)(num_dst + 1) * 4);
  steps_len_kernel<<<grid1d(num_dst + 1), 256, 0, st>>>(s, num_dst, lens);
  IGN_CHECK_LAUNCH("steps_len");
  int rc = exclusive_scan(lens, steps_rowptr, num_dst + 1, scan_ws, st);
  if (rc) return rc;
  if (steps && num_dst > 0) {
    steps_fill_kernel<<<grid1d(num_dst), 256, 0, st>>>(s, dst_sample, pos_off, pos_src, pos_col, num_dst,
                                                        steps_rowptr, steps);
    IGN_CHECK_LAUNCH("steps_fill");
  }
  return IGN_OK;
}

extern "C" int ign_steps_keys(const int32_t* steps, int64_t n, int src_id, int64_t n_rows, int32_t* keys,
                              void* stream) {
  IGN_REQUIRE(n >= 0 && n_rows >= 0 && src_id >= 0 && src_id < IGN_MAX_SOURCES, IGN_ERR_INVALID,
              "IGNNITION: steps_keys: bad argument");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(steps && keys, IGN_ERR_INVALID, "IGNNITION: steps_keys: null pointer");
  steps_keys_kernel<<<grid1d(n), 256, 0, ign_stream(stream)>>>(steps, n, src_id, (int)n_rows, keys);
  IGN_CHECK_LAUNCH("steps_keys");
  return IGN_OK;
}

extern "C" int ign_seq_meta(const int32_t* steps_rowptr, const int32_t* steps, const int32_t* order, int64_t num_dst,
                            int32_t* meta, void* stream) {
  IGN_REQUIRE(num_dst >= 0, IGN_ERR_INVALID, "IGNNITION: seq_meta: negative size");
  if (num_dst == 0) return IGN_OK;
  IGN_REQUIRE(steps_rowptr && steps && meta, IGN_ERR_INVALID, "IGNNITION: seq_meta: null pointer");
  seq_meta_kernel<<<grid1d(num_dst), 256, 0, ign_stream(stream)>>>(steps_rowptr, steps, order, num_dst,
                                                                    reinterpret_cast<int4*>(meta));
  IGN_CHECK_LAUNCH("seq_meta");
  return IGN_OK;
}

extern "C" int ign_seq_step_plan(const int32_t* meta, const int32_t* steps, int64_t num_dst, int max_steps,
                                 int32_t* nt, int32_t* off, int32_t* steps_T, void* stream) {
  IGN_REQUIRE(num_dst >= 0 && max_steps >= 1 && max_steps <= 1024, IGN_ERR_INVALID,
              "IGNNITION: seq_step_plan: bad argument (1 <= max_steps <= 1024)");
  IGN_REQUIRE(meta && steps && nt && off && steps_T, IGN_ERR_INVALID, "IGNNITION: seq_step_plan: null pointer");
  cudaStream_t st = ign_stream(stream);
  // nt doubles as the histogram scratch: cnt lives in off[] until the scan overwrites both in order
  IGN_CUDA(cudaMemsetAsync(off, 0, (size_t)(max_steps + 1) * sizeof(int), st));
  if (num_dst > 0) {
    seq_len_hist_kernel<<<grid1d(num_dst), 256, 0, st>>>(reinterpret_cast<const int4*>(meta), num_dst, max_steps, off);
    IGN_CHECK_LAUNCH("seq_len_hist");
  }
  // scan reads cnt from a copy: move the histogram into steps_T's first max_steps+1 ints first
  IGN_CUDA(cudaMemcpyAsync(steps_T, off, (size_t)(max_steps + 1) * sizeof(int), cudaMemcpyDeviceToDevice, st));
  seq_plan_scan_kernel<<<1, 32, 0, st>>>(steps_T, max_steps, nt, off);
  IGN_CHECK_LAUNCH("seq_plan_scan");
  if (num_dst > 0) {
    seq_steps_transpose_kernel<<<grid1d(num_dst), 256, 0, st>>>(reinterpret_cast<const int4*>(meta), steps, num_dst,
                                                                  max_steps, off, steps_T);
    IGN_CHECK_LAUNCH("seq_steps_transpose");
  }
  return IGN_OK;
}
