// Peer-mapped state buffers and the small kernels of the destination-partitioned graph (SURVEY.md
// section 8e): one process per GPU, every rank owns a contiguous range of destination rows and the
// authoritative state of those nodes; the full source-state array of a rank is a cudaMalloc'ed buffer
// that every peer maps through CUDA IPC, so the kernel that computes a tile of new states stores it
// into all of them (agg_gru_tc.cu) and no separate all-gather runs.
//
// This file holds the only allocations of the library: an IPC handle covers a whole cudaMalloc
// allocation, so the exchanged buffers cannot be carved out of the caller's (PyTorch's) pool.

#include "common.cuh"

namespace {

// owner[i] = rank r with bounds[r] <= dst[i] < bounds[r + 1] (bounds ascending, bounds[world] = N)
struct Bounds {
  int v[IGN_MAX_PEERS + 1];
  int world;
};
__global__ void edge_owner_kernel(const int* __restrict__ dst, int64_t n, Bounds b, int* __restrict__ owner) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = dst[i];
  int r = 0;
#pragma unroll
  for (int k = 1; k < IGN_MAX_PEERS; ++k) r += (k < b.world && d >= b.v[k]) ? 1 : 0;
  owner[i] = r;
}

// out[i] = in[perm[i]] + add
__global__ void gather_int_kernel(const int* __restrict__ in, const int* __restrict__ perm, int64_t n, int add,
                                  int* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[perm ? perm[i] : i] + add;
}

// flags[col[e]] = 1 for every slot
__global__ void mark_rows_kernel(const int* __restrict__ col, int64_t n, int* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int c = col[i];
    if (c >= 0) flags[c] = 1;
  }
}

// dst[rows[i]] = src[rows[i]] for the listed rows of a [*, width] fp32 array; G = width / 4 lanes per row, so
// a warp writes whole rows (peer stores over NVLink leave as full 128-byte segments)
__global__ void rows_put_kernel(const float* __restrict__ src, const int* __restrict__ rows, int64_t n, int width,
                                float* __restrict__ dst) {
  const int q = width / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = n * q;
  for (int64_t j = i; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = j / q;
    const int c = (int)(j - k * q) * 4;
    const int64_t r = rows[k];
    st_f4(dst + r * width + c, ld_stream_f4(src + r * width + c));
  }
}

// dst[rows[i], :] = packed[i, :]: the receiving side of a packed boundary exchange
__global__ void rows_unpack_kernel(const float* __restrict__ packed, const int* __restrict__ rows, int64_t n, int width,
                                   float* __restrict__ dst) {
  const int q = width / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = n * q;
  for (int64_t j = i; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = j / q;
    const int c = (int)(j - k * q) * 4;
    st_f4(dst + (int64_t)rows[k] * width + c, ld_stream_f4(packed + k * width + c));
  }
}

inline unsigned grid_for(int64_t n) { return (unsigned)(n > 0 ? ign_cdiv(n, 256) : 1); }

}  // namespace

extern "C" int ign_peer_alloc(size_t bytes, void** ptr) {
  IGN_REQUIRE(ptr && bytes > 0, IGN_ERR_INVALID, "IGNNITION: peer_alloc: bad argument");
  IGN_CUDA(cudaMalloc(ptr, bytes));
  return IGN_OK;
}

extern "C" int ign_peer_free(void* ptr) {
  if (ptr) IGN_CUDA(cudaFree(ptr));
  return IGN_OK;
}

extern "C" int ign_peer_export(void* ptr, void* handle64) {
  IGN_REQUIRE(ptr && handle64, IGN_ERR_INVALID, "IGNNITION: peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == IGN_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t hd;
  IGN_CUDA(cudaIpcGetMemHandle(&hd, ptr));
  memcpy(handle64, &hd, sizeof(hd));
  return IGN_OK;
}

extern "C" int ign_peer_open(const void* handle64, void** ptr) {
  IGN_REQUIRE(ptr && handle64, IGN_ERR_INVALID, "IGNNITION: peer_open: null pointer");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, sizeof(hd));
  IGN_CUDA(cudaIpcOpenMemHandle(ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  return IGN_OK;
}

extern "C" int ign_peer_close(void* ptr) {
  if (ptr) IGN_CUDA(cudaIpcCloseMemHandle(ptr));
  return IGN_OK;
}

extern "C" int ign_edge_owner(const int32_t* dst, int64_t n_edges, const int32_t* bounds, int world, int32_t* owner,
                              void* stream) {
  IGN_REQUIRE(n_edges >= 0 && world >= 1 && world <= IGN_MAX_PEERS && bounds, IGN_ERR_INVALID,
              "IGNNITION: edge_owner: between 1 and %d ranks", IGN_MAX_PEERS);
  if (n_edges == 0) return IGN_OK;
  IGN_REQUIRE(dst && owner, IGN_ERR_INVALID, "IGNNITION: edge_owner: null pointer");
  Bounds b;
  b.world = world;
  for (int k = 0; k <= IGN_MAX_PEERS; ++k) b.v[k] = bounds[k <= world ? k : world];
  edge_owner_kernel<<<grid_for(n_edges), 256, 0, ign_stream(stream)>>>(dst, n_edges, b, owner);
  IGN_CHECK_LAUNCH("edge_owner");
  return IGN_OK;
}

extern "C" int ign_gather_int(const int32_t* in, const int32_t* perm, int64_t n, int add, int32_t* out, void* stream) {
  IGN_REQUIRE(n >= 0, IGN_ERR_INVALID, "IGNNITION: gather_int: negative size");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(in && out, IGN_ERR_INVALID, "IGNNITION: gather_int: null pointer");
  gather_int_kernel<<<grid_for(n), 256, 0, ign_stream(stream)>>>(in, perm, n, add, out);
  IGN_CHECK_LAUNCH("gather_int");
  return IGN_OK;
}

extern "C" int ign_mark_rows(const int32_t* col, int64_t n, int32_t* flags, void* stream) {
  IGN_REQUIRE(n >= 0, IGN_ERR_INVALID, "IGNNITION: mark_rows: negative size");
  if (n == 0) return IGN_OK;
  IGN_REQUIRE(col && flags, IGN_ERR_INVALID, "IGNNITION: mark_rows: null pointer");
  mark_rows_kernel<<<grid_for(n), 256, 0, ign_stream(stream)>>>(col, n, flags);
  IGN_CHECK_LAUNCH("mark_rows");
  return IGN_OK;
}

extern "C" int ign_rows_put(const float* src, const int32_t* rows, int64_t n_rows, int width, float* dst, void* stream) {
  IGN_REQUIRE(n_rows >= 0 && width > 0 && width % 4 == 0, IGN_ERR_INVALID, "IGNNITION: rows_put: bad shape");
  if (n_rows == 0) return IGN_OK;
  IGN_REQUIRE(src && rows && dst, IGN_ERR_INVALID, "IGNNITION: rows_put: null pointer");
  const int64_t total = n_rows * (width / 4);
  int64_t blocks = ign_cdiv(total, 256);
  if (blocks > IGN_NUM_SMS * 16) blocks = IGN_NUM_SMS * 16;
  rows_put_kernel<<<(unsigned)blocks, 256, 0, ign_stream(stream)>>>(src, rows, n_rows, width, dst);
  IGN_CHECK_LAUNCH("rows_put");
  return IGN_OK;
}

extern "C" int ign_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return IGN_OK;
  IGN_REQUIRE(dst && src, IGN_ERR_INVALID, "IGNNITION: peer_copy: null pointer");
  IGN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ign_stream(stream)));
  return IGN_OK;
}

extern "C" int ign_rows_unpack(const float* packed, const int32_t* rows, int64_t n_rows, int width, float* dst,
                               void* stream) {
  IGN_REQUIRE(n_rows >= 0 && width > 0 && width % 4 == 0, IGN_ERR_INVALID, "IGNNITION: rows_unpack: bad shape");
  if (n_rows == 0) return IGN_OK;
  IGN_REQUIRE(packed && rows && dst, IGN_ERR_INVALID, "IGNNITION: rows_unpack: null pointer");
  const int64_t total = n_rows * (width / 4);
  int64_t blocks = ign_cdiv(total, 256);
  if (blocks > IGN_NUM_SMS * 16) blocks = IGN_NUM_SMS * 16;
  rows_unpack_kernel<<<(unsigned)blocks, 256, 0, ign_stream(stream)>>>(packed, rows, n_rows, width, dst);
  IGN_CHECK_LAUNCH("rows_unpack");
  return IGN_OK;
}
