// Peer-store rate probe (one process, all visible GPUs): how fast can a kernel on GPU a push 32 KB tiles that sit
// in shared memory into the memory of its peers over NVLink, by store method?  Motivates the exchange path of
// agg_gru_tc.cu (profiles/r2_p2p_rate.md).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/p2p_rate tools/p2p_rate.cu
//   tools/p2p_rate [MB per peer]
//
// Methods (every CTA owns tiles t = blockIdx.x, + gridDim.x, ...; a tile is 32 KB, contiguous in the destination):
//   0  bulk      one cp.async.bulk.global.shared::cta per tile and peer, wait_group.read 0 before the tile is reused
//   1  bulk x4   the tile cut into four 8 KB bulk stores
//   2  st.v4     128 threads store the tile with st.global.v4 (a warp writes 512 contiguous bytes per instruction)
//   3  st.v4 row thread = one 256-byte row, 16 stores of 16 bytes each (what a thread-per-row epilogue would do)
//   4  bulk, two staging tiles (wait_group.read 1)
// Patterns: one direction (GPU 0 -> GPU 1), both directions at once, and all-to-all (every GPU -> every peer).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                                       \
  do {                                                                                              \
    cudaError_t e_ = (x);                                                                           \
    if (e_ != cudaSuccess) {                                                                        \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));            \
      exit(1);                                                                                      \
    }                                                                                               \
  } while (0)

constexpr int TILE = 32768;
constexpr int MAXP = 8;
struct Peers {
  float* p[MAXP];
  int n;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int METHOD>
__global__ void __launch_bounds__(128, 1) push_kernel(Peers peers, int64_t ntiles, int row0_tiles) {
  extern __shared__ __align__(1024) unsigned char stage[];   // 1 or 2 tiles
  const int tid = threadIdx.x;
  // fill the staging tile(s) once (generic-proxy writes, then the async proxy may read them)
  for (int i = tid; i < (METHOD == 4 ? 2 : 1) * TILE / 16; i += 128)
    reinterpret_cast<float4*>(stage)[i] = make_float4(1.f, 2.f, 3.f, (float)blockIdx.x);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  int k = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++k) {
    const int64_t off = (int64_t)(row0_tiles + t) * (TILE / 4);
    if (METHOD == 0 || METHOD == 1 || METHOD == 4) {
      if (tid == 0) {
        const unsigned char* s = stage + (METHOD == 4 ? (k & 1) * TILE : 0);
        if (METHOD == 4) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        for (int q = 0; q < peers.n; ++q) {
          constexpr int PIECES = METHOD == 1 ? 4 : 1;
#pragma unroll
          for (int c = 0; c < PIECES; ++c)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                             peers.p[q] + off + c * (TILE / 4 / PIECES)),
                         "r"(smem_u32(s + c * (TILE / PIECES))), "r"(TILE / PIECES)
                         : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else if (METHOD == 2) {
      for (int q = 0; q < peers.n; ++q) {
        float4* d = reinterpret_cast<float4*>(peers.p[q] + off);
#pragma unroll 4
        for (int i = tid; i < TILE / 16; i += 128) d[i] = reinterpret_cast<const float4*>(stage)[i];
      }
    } else {
      for (int q = 0; q < peers.n; ++q) {
        float4* d = reinterpret_cast<float4*>(peers.p[q] + off) + tid * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) d[i] = reinterpret_cast<const float4*>(stage)[tid * 16 + i];
      }
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}


// HBM hog: every thread streams 16-byte pieces of pseudo-random 256-byte rows of a large table (the gather of
// agg_gru_tc.cu), so that the peer stores above can be timed while the memory system of the sender, the receiver
// or both is as busy as it is in the real kernel.
__global__ void __launch_bounds__(256) hog_kernel(const float4* __restrict__ table, int64_t rows, int iters,
                                                  float4* __restrict__ sink) {
  const int lane16 = threadIdx.x & 15;
  uint32_t s = (blockIdx.x * 256 + threadIdx.x) / 16 * 2654435761u + 12345u;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i < iters; ++i) {
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s = s * 1664525u + 1013904223u;
      const int64_t r = (int64_t)(s >> 4) % rows;
      v[j] = __ldg(table + r * 16 + lane16);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
  }
  if (acc.x == 123.456f) sink[0] = acc;
}

typedef void (*Kern)(Peers, int64_t, int);
static Kern kerns[5] = {push_kernel<0>, push_kernel<1>, push_kernel<2>, push_kernel<3>, push_kernel<4>};
static const char* names[5] = {"bulk 32K", "bulk 4x8K", "st.v4 coalesced", "st.v4 thread-per-row", "bulk 32K x2 stages"};

int main(int argc, char** argv) {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev > MAXP) ndev = MAXP;
  const int64_t mb = argc > 1 ? atoll(argv[1]) : 320;          // per (sender, receiver) pair
  const int64_t ntiles = mb * 1024 * 1024 / TILE;
  printf("devices %d, %lld MB per pair, tile %d B\n", ndev, (long long)mb, TILE);
  if (ndev < 2) { printf("need two GPUs\n"); return 0; }
  std::vector<float*> buf(ndev);
  std::vector<cudaStream_t> st(ndev);
  std::vector<cudaEvent_t> e0(ndev), e1(ndev);
  for (int d = 0; d < ndev; ++d) {
    CK(cudaSetDevice(d));
    for (int p = 0; p < ndev; ++p)
      if (p != d) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, d, p));
        if (!can) { printf("no peer access %d -> %d\n", d, p); return 0; }
        cudaError_t e = cudaDeviceEnablePeerAccess(p, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        cudaGetLastError();
      }
    CK(cudaMalloc(&buf[d], (size_t)ndev * ntiles * TILE));     // slot s of the buffer receives from GPU s
    CK(cudaMemset(buf[d], 0, (size_t)ndev * ntiles * TILE));
    CK(cudaStreamCreate(&st[d]));
    CK(cudaEventCreate(&e0[d]));
    CK(cudaEventCreate(&e1[d]));
    for (int m = 0; m < 5; ++m)
      CK(cudaFuncSetAttribute(kerns[m], cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE));
  }
  // pattern 0: 0 -> 1; 1: 0 <-> 1; 2: all-to-all (only when more than two GPUs)
  for (int pattern = 0; pattern < (ndev > 2 ? 3 : 2); ++pattern) {
    for (int m = 0; m < 5; ++m) {
      for (int grid : {148, 296}) {
        if (grid == 296 && m != 2 && m != 0) continue;
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
          const int senders = pattern == 0 ? 1 : pattern == 1 ? 2 : ndev;
          for (int d = 0; d < senders; ++d) {
            CK(cudaSetDevice(d));
            Peers pr;
            pr.n = 0;
            for (int p = 0; p < (pattern == 2 ? ndev : 2); ++p)
              if (p != d) pr.p[pr.n++] = buf[p];
            CK(cudaEventRecord(e0[d], st[d]));
            // two CTAs per SM need two kernels' worth of shared memory: 296 CTAs of 32 KB fit
            kerns[m]<<<grid, 128, (m == 4 ? 2 : 1) * TILE, st[d]>>>(pr, ntiles, (int)(d * ntiles));
            CK(cudaEventRecord(e1[d], st[d]));
          }
          float worst = 0;
          for (int d = 0; d < senders; ++d) {
            CK(cudaSetDevice(d));
            CK(cudaStreamSynchronize(st[d]));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0[d], e1[d]));
            if (ms > worst) worst = ms;
          }
          if (rep > 0 && worst < best) best = worst;
        }
        const int npeers = pattern == 2 ? ndev - 1 : 1;
        const double gb = (double)npeers * ntiles * TILE / 1e9;
        printf("pattern %d  %-22s grid %3d  %8.3f ms  %7.1f GB/s out per GPU\n", pattern, names[m], grid, best,
               gb / (best / 1e3));
      }
    }
  }

  // ---- the same pushes while HBM hogs run: on the receiver, on the sender, on both (both directions pushed)
  {
    const int64_t rows = (int64_t)10 * 1000 * 1000;              // 2.56 GB table of 256-byte rows
    std::vector<float4*> table(2), sink(2);
    std::vector<cudaStream_t> hs(2);
    std::vector<cudaEvent_t> h0(2), h1(2);
    for (int d = 0; d < 2; ++d) {
      CK(cudaSetDevice(d));
      CK(cudaMalloc(&table[d], rows * 256));
      CK(cudaMemset(table[d], 0, rows * 256));
      CK(cudaMalloc(&sink[d], 256));
      CK(cudaStreamCreateWithFlags(&hs[d], cudaStreamNonBlocking));
      CK(cudaEventCreate(&h0[d]));
      CK(cudaEventCreate(&h1[d]));
    }
    const int hog_grid = 148 * 6, hog_iters = 1400;              // ~ 148*6*256*8*16 B * iters = 40 GB, ~6-8 ms
    const double hog_gb = (double)hog_grid * 256 * 8 * 16 * hog_iters / 1e9;
    // hog_on: bit 0 = GPU 0 (the sender in one-direction runs), bit 1 = GPU 1
    for (int both_dirs = 0; both_dirs < 1; ++both_dirs)
      for (int hog_on : {0, 1})
        for (int m : {0, 2, 4}) {
          float push_ms[2] = {0, 0}, hog_ms[2] = {0, 0};
          for (int rep = 0; rep < 3; ++rep) {
            for (int d = 0; d < 2; ++d)
              if (hog_on & (1 << d)) {
                CK(cudaSetDevice(d));
                CK(cudaEventRecord(h0[d], hs[d]));
                hog_kernel<<<hog_grid, 256, 0, hs[d]>>>(table[d], rows, hog_iters, sink[d]);
                CK(cudaEventRecord(h1[d], hs[d]));
              }
            for (int d = 0; d < 1 + both_dirs; ++d) {
              CK(cudaSetDevice(d));
              Peers pr;
              pr.n = 1;
              pr.p[0] = buf[1 - d];
              CK(cudaEventRecord(e0[d], st[d]));
              kerns[m]<<<148, 128, (m == 4 ? 2 : 1) * TILE, st[d]>>>(pr, ntiles, (int)(d * ntiles));
              CK(cudaEventRecord(e1[d], st[d]));
            }
            for (int d = 0; d < 2; ++d) {
              CK(cudaSetDevice(d));
              CK(cudaDeviceSynchronize());
              if (d < 1 + both_dirs) CK(cudaEventElapsedTime(&push_ms[d], e0[d], e1[d]));
              if (hog_on & (1 << d)) CK(cudaEventElapsedTime(&hog_ms[d], h0[d], h1[d]));
            }
          }
          const double gb = (double)ntiles * TILE / 1e9;
          printf("loaded  dirs %d  hog on gpu mask %d  %-20s push0 %7.3f ms %6.1f GB/s", both_dirs + 1, hog_on, names[m],
                 push_ms[0], gb / (push_ms[0] / 1e3));
          if (both_dirs) printf("  push1 %7.3f ms %6.1f GB/s", push_ms[1], gb / (push_ms[1] / 1e3));
          for (int d = 0; d < 2; ++d)
            if (hog_on & (1 << d)) printf("  hog%d %6.2f ms %6.0f GB/s", d, hog_ms[d], hog_gb / (hog_ms[d] / 1e3));
          printf("\n");
        }

    // ---- sender-side load only, by how much the hog keeps in flight, with the copy engine as a fourth method
    {
      float* src0 = nullptr;
      CK(cudaSetDevice(0));
      CK(cudaMalloc(&src0, (size_t)ntiles * TILE));
      CK(cudaMemset(src0, 0, (size_t)ntiles * TILE));
      for (int g : {1, 2, 4, 6})
        for (int m : {0, 2, 12, 99}) {                           // 12: st.v4 from four CTAs per SM; 99: cudaMemcpyPeerAsync
          float push_ms = 0, hog_ms = 0;
          const int iters = hog_iters * 6 / g;
          for (int rep = 0; rep < 3; ++rep) {
            CK(cudaSetDevice(0));
            CK(cudaEventRecord(h0[0], hs[0]));
            hog_kernel<<<148 * g, 256, 0, hs[0]>>>(table[0], rows, iters, sink[0]);
            CK(cudaEventRecord(h1[0], hs[0]));
            Peers pr;
            pr.n = 1;
            pr.p[0] = buf[1];
            CK(cudaEventRecord(e0[0], st[0]));
            if (m == 99) CK(cudaMemcpyPeerAsync(buf[1], 1, src0, 0, (size_t)ntiles * TILE, st[0]));
            else if (m == 12) kerns[2]<<<148 * 4, 128, TILE, st[0]>>>(pr, ntiles, 0);
            else kerns[m]<<<148, 128, TILE, st[0]>>>(pr, ntiles, 0);
            CK(cudaEventRecord(e1[0], st[0]));
            CK(cudaDeviceSynchronize());
            CK(cudaEventElapsedTime(&push_ms, e0[0], e1[0]));
            CK(cudaEventElapsedTime(&hog_ms, h0[0], h1[0]));
          }
          const double gb = (double)ntiles * TILE / 1e9;
          const double hgb = (double)148 * g * 256 * 8 * 16 * iters / 1e9;
          printf("sender hog %d x 256 thr/SM (%3d KB in flight/SM)  %-18s push %7.3f ms %6.1f GB/s   hog %6.2f ms %6.0f GB/s\n",
                 g, g * 256 * 8 * 16 / 1024, m == 99 ? "copy engine" : m == 12 ? "st.v4 4 CTAs/SM" : names[m], push_ms,
                 gb / (push_ms / 1e3), hog_ms, hgb / (hog_ms / 1e3));
        }
    }
  }
  // check: slot 0 of GPU 1 holds the pattern
  CK(cudaSetDevice(1));
  float v[4];
  CK(cudaMemcpy(v, buf[1], 16, cudaMemcpyDeviceToHost));
  printf("check %.0f %.0f %.0f\n", v[0], v[1], v[2]);
  return 0;
}
