// Micro-benchmark: HBM write bandwidth of the training forward's image-store PATTERN, without any compute.
// 148 persistent CTAs; CTA c takes tiles c, c + 148, ...; per tile it writes 14 "layer images" of 64 KB each.
//   mode 0: layer-major images (the current T16 workspace): image l holds all tiles, CTA writes 64 KB at img[l] + tile * 64 KB
//   mode 1: tile-major: the 14 x 64 KB of a tile are contiguous
//   mode 2: like 0 but every warp writes its 4 KB piece as 8 x 512 B (st.global.v4 per lane) in the kernel's (row-group, chunk) order
// nvcc -gencode arch=compute_100a,code=sm_100a scripts/micro/store_pattern.cu -o build/store_pattern
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kLayers = 14;
constexpr size_t kTileBytes = 65536;

__global__ void __launch_bounds__(512, 1) k(uint8_t* base, int n_tiles, int mode) {
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int l = 0; l < kLayers; ++l) {
      uint8_t* dst = mode == 1 ? base + ((size_t)tile * kLayers + l) * kTileBytes
                               : base + (size_t)l * n_tiles * kTileBytes + (size_t)tile * kTileBytes;
      if (mode == 2) {
        // warp w (16 warps): quarter q = w & 3 rows 32q..32q+31, 64 columns [64 (w >> 2), +64): 8 chunks of 16 B per lane
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, q = w & 3, cb = (w >> 2) * 8;
        const int row = q * 32 + lane;
        uint8_t* p = dst + (size_t)(row >> 4) * 8192 + (size_t)(row & 15) * 16 + (size_t)cb * 256;
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(p + c * 256) = v;
      } else {
        for (int i = threadIdx.x; i < (int)(kTileBytes / 16); i += blockDim.x) reinterpret_cast<uint4*>(dst)[i] = v;
      }
    }
  }
}

int main() {
  const int n_tiles = 6144;
  uint8_t* buf;
  const size_t bytes = (size_t)n_tiles * kLayers * kTileBytes;
  cudaMalloc(&buf, bytes);
  for (int grid : {148, 296}) {
    for (int mode = 0; mode < 3; ++mode) {
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<grid, 512>>>(buf, n_tiles, mode);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
      }
      printf("grid %d mode %d: %.3f ms, %.0f GB/s\n", grid, mode, best, bytes / (best * 1e-3) / 1e9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
