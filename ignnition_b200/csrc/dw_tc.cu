// Weight gradient of a Dense layer on the tcgen05 tensor cores (sm_100a):
//     dW[k, n] += X^T[k, m] . dZ[m, n]          (tf.gradients through Dense, generate_model.py:791)
// a tall-skinny product whose reduction runs over the m rows of the batch (2.26 M paths for the
// RouteNet readout), k, n <= 256.
//
// Both operands arrive row-major with the REDUCTION index outermost, i.e. "MN-major" in UMMA terms, so
// no transpose exists anywhere: a [16 rows][32 floats] block is stored as 128-byte rows (coalesced
// float4 loads, hi / lo split, 16-byte stores) and the instruction descriptor sets the two "MN-major"
// bits.  The one shared-memory layout tcgen05 accepts for MN-major tf32 operands is
// SWIZZLE_128B_BASE32B: atoms of 4 reduction rows x 128 bytes whose 32-byte chunks are XOR-ed with
// (row & 3) -- the plain SWIZZLE_128B image of the forward kernels makes the instruction a silent
// no-op (measured: accumulators stay zero).  One tcgen05.mma (K = 8 for tf32) consumes 8 rows = two
// atoms (SBO = 512 bytes) of every block; feature blocks of 32 sit LBO = one image apart.
//
// 3xTF32: X = X_hi + X_lo, dZ = dZ_hi + dZ_lo.  For k <= 64 the M = 128 lanes of one instruction hold
// [X_hi | X_lo] side by side (the A descriptor simply spans both images), so 2 instructions per K
// step give all four partial products and the lanes are added when the accumulator is flushed; for
// k = 128 / 256 the usual three products are issued per 128-feature tile.
//
// Persistent CTAs split the m rows; 16 producer warps keep two stages of loads in flight, one warp
// issues, accumulators stay in TMEM for the CTA's whole life and are flushed once with atomics.

#include "tc_common.cuh"

using namespace ign_tc;

namespace {

constexpr int R = 16;                      // rows per stage = 2 K-steps
constexpr int IMGB = R * 128;              // bytes of one [16 x 32] fp32 image
constexpr int PROD_WARPS = 16;
constexpr int MMA_WARP = PROD_WARPS;
constexpr int DW_THREADS = 32 * (PROD_WARPS + 1);
constexpr int PROD_THREADS = 32 * PROD_WARPS;

// shared-memory descriptor of an MN-major SWIZZLE_128B_BASE32B operand: 32 MN elements per 128-byte row,
// the next block of 32 MN elements lbo bytes further, 4 reduction rows per 512-byte atom
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
// 16-byte chunk c4 of reduction row r as hi / lo into two MN-major images
__device__ __forceinline__ void store_split_mn(unsigned char* img_hi, unsigned char* img_lo, int r, int c4, float4 v) {
  float4 hi, lo;
  tf32_split(v.x, hi.x, lo.x);
  tf32_split(v.y, hi.y, lo.y);
  tf32_split(v.z, hi.z, lo.z);
  tf32_split(v.w, hi.w, lo.w);
  const int off = r * 128 + ((((c4 >> 1) ^ (r & 3)) << 5) | ((c4 & 1) << 4));
  *reinterpret_cast<float4*>(img_hi + off) = hi;
  *reinterpret_cast<float4*>(img_lo + off) = lo;
}
// kind::tf32, D fp32, A and B MN-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_mn(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

template <int KB, int NB>
struct Cfg {
  static constexpr int K = 32 * KB, N = 32 * NB;
  static constexpr int BLOCKS = 2 * (KB + NB);           // A_hi[KB] A_lo[KB] B_hi[NB] B_lo[NB]
  static constexpr int STAGE = BLOCKS * IMGB;
  static constexpr int STAGES = (200 * 1024 / STAGE) < 2 ? 2 : ((200 * 1024 / STAGE) > 6 ? 6 : (200 * 1024 / STAGE));
  static constexpr bool LANES = KB <= 2;                 // hi and lo of X share one instruction's lanes
  static constexpr int MT = LANES ? 1 : KB / 4;          // 128-lane accumulator tiles
  static constexpr int COLS = MT * N;
  static constexpr int TMEM_COLS = COLS <= 32 ? 32 : COLS <= 64 ? 64 : COLS <= 128 ? 128 : COLS <= 256 ? 256 : 512;
  static constexpr int F4 = R * 8 * (KB + NB);           // float4 loads per stage
  static constexpr int PER = (F4 + PROD_THREADS - 1) / PROD_THREADS;
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE;
  static_assert(BLOCKS >= 4, "the M = 128 descriptor of the lane layout spans four images");
  static_assert(COLS <= 512, "accumulators must fit in TMEM");
};

template <int KB, int NB>
__global__ void __launch_bounds__(DW_THREADS, 1) dw_tc_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                              int64_t m, float* __restrict__ dw) {
  using C = Cfg<KB, NB>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full[C::STAGES], bar_free[C::STAGES], bar_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&bar_full[s], PROD_WARPS);
      mbar_init(&bar_free[s], 1);
    }
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) tmem_alloc(&tmem_base_s, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // this CTA's 16-row chunks
  const int64_t nchunks = (m + R - 1) / R;
  const int64_t per = (nchunks + gridDim.x - 1) / gridDim.x;
  const int64_t c0 = (int64_t)blockIdx.x * per;
  const int64_t c1 = c0 + per < nchunks ? c0 + per : nchunks;
  const int64_t n_my = c1 > c0 ? c1 - c0 : 0;

  if (warp < PROD_WARPS) {
    // ================================ producers ================================
    float4 cur[C::PER], nxt[C::PER];
    auto load = [&](int64_t chunk, float4 (&v)[C::PER]) {
      const int64_t row0 = chunk * R;
#pragma unroll
      for (int p = 0; p < C::PER; ++p) {
        const int idx = tid + p * PROD_THREADS;
        v[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < R * 8 * KB) {                             // X part: row-major [R][K]
          const int r = idx / (8 * KB), w = idx % (8 * KB);
          if (row0 + r < m) v[p] = ld_stream_f4(x + (row0 + r) * C::K + w * 4);
        } else if (idx < C::F4) {                           // dZ part: row-major [R][N]
          const int j = idx - R * 8 * KB;
          const int r = j / (8 * NB), w = j % (8 * NB);
          if (row0 + r < m) v[p] = ld_stream_f4(dz + (row0 + r) * C::N + w * 4);
        }
      }
    };
    auto store = [&](unsigned char* st, const float4 (&v)[C::PER]) {
#pragma unroll
      for (int p = 0; p < C::PER; ++p) {
        const int idx = tid + p * PROD_THREADS;
        if (idx < R * 8 * KB) {
          const int r = idx / (8 * KB), w = idx % (8 * KB);
          unsigned char* hi = st + (w >> 3) * IMGB;
          store_split_mn(hi, hi + KB * IMGB, r, w & 7, v[p]);
        } else if (idx < C::F4) {
          const int j = idx - R * 8 * KB;
          const int r = j / (8 * NB), w = j % (8 * NB);
          unsigned char* hi = st + (2 * KB + (w >> 3)) * IMGB;
          store_split_mn(hi, hi + NB * IMGB, r, w & 7, v[p]);
        }
      }
    };
    if (n_my > 0) load(c0, cur);
    for (int64_t i = 0; i < n_my; ++i) {
      const int s = (int)(i % C::STAGES);
      if (i + 1 < n_my) load(c0 + i + 1, nxt);
      if (i >= C::STAGES) mbar_wait(&bar_free[s], (uint32_t)((i / C::STAGES) - 1) & 1);
      store(smem + (size_t)s * C::STAGE, cur);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[s]);
#pragma unroll
      for (int p = 0; p < C::PER; ++p) cur[p] = nxt[p];
    }
    // ================================ flush ================================
    if (n_my > 0) {
      mbar_wait(&bar_done, 0);
      tc_fence_after();
      const int lg = warp & 3, cq = warp >> 2;
      const int L = lg * 32 + lane;                         // TMEM lane = feature (lane layout: hi | lo | unused)
      bool used = true;
      int f = L;
      if (C::LANES) {
        used = L < 2 * C::K;
        f = L < C::K ? L : L - C::K;
      }
      constexpr int CPW = C::COLS / 4;                      // accumulator columns per warp quarter
      if (used) {                                           // warp-uniform: K is a multiple of 32
#pragma unroll 1
        for (int cb = 0; cb < CPW; cb += 8) {
          const int col = cq * CPW + cb;
          uint32_t v[8];
          tmem_ld8_nowait(tmem_base + ((uint32_t)(lg * 32) << 16) + col, v);
          tmem_ld_wait();
          const int mt = col / C::N, n0 = col % C::N;
          float* dst = dw + (int64_t)(mt * 128 + f) * C::N + n0;
          red_add_v4(dst, __uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
          red_add_v4(dst + 4, __uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
        }
      }
    }
  } else {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc = umma_idesc_mn(C::N);
    for (int64_t i = 0; i < n_my; ++i) {
      const int s = (int)(i % C::STAGES);
      mbar_wait(&bar_full[s], (uint32_t)(i / C::STAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t st = smem_u32(smem + (size_t)s * C::STAGE);
        const uint32_t b_hi = st + 2 * KB * IMGB, b_lo = b_hi + NB * IMGB;
#pragma unroll
        for (int ks = 0; ks < R / 8; ++ks) {
          const uint32_t ko = ks * 1024;
          const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
          if (C::LANES) {
            umma_tf32(tmem_base, umma_desc_mn(st + ko, IMGB), umma_desc_mn(b_hi + ko, IMGB), idesc, acc);
            umma_tf32(tmem_base, umma_desc_mn(st + ko, IMGB), umma_desc_mn(b_lo + ko, IMGB), idesc, 1u);
          } else {
#pragma unroll
            for (int mt = 0; mt < C::MT; ++mt) {
              const uint32_t a_hi = st + mt * 4 * IMGB, a_lo = a_hi + KB * IMGB;
              const uint32_t d = tmem_base + mt * C::N;
              umma_tf32(d, umma_desc_mn(a_hi + ko, IMGB), umma_desc_mn(b_hi + ko, IMGB), idesc, acc);
              umma_tf32(d, umma_desc_mn(a_lo + ko, IMGB), umma_desc_mn(b_hi + ko, IMGB), idesc, 1u);
              umma_tf32(d, umma_desc_mn(a_hi + ko, IMGB), umma_desc_mn(b_lo + ko, IMGB), idesc, 1u);
            }
          }
        }
        umma_commit(&bar_free[s]);
      }
      __syncwarp();
    }
    if (lane == 0 && n_my > 0) umma_commit(&bar_done);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int KB, int NB>
int launch(const float* x, const float* dz, int64_t m, float* dw, cudaStream_t st) {
  using C = Cfg<KB, NB>;
  if (IGN_ONCE_PER_DEVICE()) {
    IGN_CUDA(cudaFuncSetAttribute(dw_tc_kernel<KB, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  }
  int sms = IGN_NUM_SMS, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // at least 64 chunks (1024 rows) per CTA so that the atomic flush stays a small part of the work
  const int64_t nchunks = ign_cdiv(m, R);
  int64_t grid = ign_cdiv(nchunks, 64);
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  dw_tc_kernel<KB, NB><<<(unsigned)grid, DW_THREADS, C::SMEM, st>>>(x, dz, m, dw);
  IGN_CHECK_LAUNCH("dense_bwd_dw_tc");
  return IGN_OK;
}

}  // namespace

bool ign_dw_tc_supported(int k, int n) {
  return (k == 32 || k == 64 || k == 128 || k == 256) && (n == 32 || n == 64 || n == 128 || n == 256);
}

#define IGN_DW_N(KB)                                                  \
  switch (n / 32) {                                                   \
    case 1: return launch<KB, 1>(x, dz, m, dw, st);                   \
    case 2: return launch<KB, 2>(x, dz, m, dw, st);                   \
    case 4: return launch<KB, 4>(x, dz, m, dw, st);                   \
    case 8: return launch<KB, 8>(x, dz, m, dw, st);                   \
    default: break;                                                   \
  }

// dw[k, n] += x^T[k, m] dz[m, n]; shapes: ign_dw_tc_supported and n in {32, 64, 128, 256}
int ign_dw_tc_launch(const float* x, const float* dz, int64_t m, int k, int n, float* dw, cudaStream_t st) {
  switch (k / 32) {
    case 1: IGN_DW_N(1) break;
    case 2: IGN_DW_N(2) break;
    case 4: IGN_DW_N(4) break;
    case 8: IGN_DW_N(8) break;
    default: break;
  }
  ign_set_error("IGNNITION: dense_bwd: tensor-core weight gradient not built for k = %d, n = %d", k, n);
  return IGN_ERR_UNSUPPORTED;
}
