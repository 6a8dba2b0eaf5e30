// HBM write bandwidth of the GEMM epilogue's store patterns: a persistent grid (148 x W warps) writes a [T, N] bf16
// matrix tile by tile (128 x 256 tiles, warp = 32 x 32 chunk as in gemm_tcgen05.cu) with different per-instruction shapes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_pattern_bw store_pattern_bw.cu ; run: ./store_pattern_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// MODE 0: 8 B / lane, 8 lanes per row: 4 rows x 64 B per instruction (the current bf16 epilogue)
// MODE 1: 16 B / lane, 4 lanes per row: 8 rows x 64 B per instruction
// MODE 2: 16 B / lane, 8 lanes per row: 4 rows x 128 B per instruction (64-column chunks)
// MODE 3: 16 B / lane, 16 lanes per row: 2 rows x 256 B per instruction (128-column chunks)
// MODE 4: 16 B / lane, 32 lanes per row: 1 row x 512 B (whole 256-column tile row)
// MODE 5: the accumulator's native layout without a transpose: lane = row, 4 x 16 B back to back per 32 columns (32 rows x 16 B per instruction)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(uint16_t* out, int T, int N) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int n_tiles = N / 256, m_tiles = T / 128;
  for (int tile = blockIdx.x; tile < n_tiles * m_tiles; tile += gridDim.x) {
    const int nb = tile % n_tiles, mb = tile / n_tiles;
    // the tile is 128 x 256 bf16 = 64 KB; split into per-warp pieces of 32 rows x CW columns
    constexpr int CW = (MODE <= 1 || MODE == 5) ? 32 : MODE == 2 ? 64 : MODE == 3 ? 128 : 256;
    constexpr int PIECES = 4 * (256 / CW);
    for (int pc = warp; pc < PIECES; pc += nw) {
      const int q = pc & 3, c = pc >> 2;
      const int row0 = mb * 128 + q * 32, col0 = nb * 256 + c * CW;
      if (MODE == 0) {
        const int cch = lane & 7, rs = lane >> 3;
#pragma unroll
        for (int it = 0; it < 8; ++it)
          *reinterpret_cast<uint2*>(out + (size_t)(row0 + it * 4 + rs) * N + col0 + cch * 4) = make_uint2(tile, lane);
      } else if (MODE == 5) {
#pragma unroll
        for (int it = 0; it < 4; ++it)
          *reinterpret_cast<uint4*>(out + (size_t)(row0 + lane) * N + col0 + it * 8) = make_uint4(tile, lane, it, 0);
      } else {
        constexpr int LPR = CW / 8;           // lanes per row
        constexpr int RPI = 32 / LPR;         // rows per instruction
        const int cch = lane % LPR, rs = lane / LPR;
#pragma unroll
        for (int it = 0; it < 32 / RPI; ++it)
          *reinterpret_cast<uint4*>(out + (size_t)(row0 + it * RPI + rs) * N + col0 + cch * 8) = make_uint4(tile, lane, it, 0);
      }
    }
  }
}

template <int MODE>
void run(uint16_t* d, int T, int N, int threads, const char* name) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) k<MODE><<<148, threads>>>(d, T, N);
  float best = 1e9f;
  for (int i = 0; i < 8; ++i) {
    cudaEventRecord(a); k<MODE><<<148, threads>>>(d, T, N); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  printf("%-44s N=%4d threads=%3d  %7.1f us  %7.1f GB/s\n", name, N, threads, best * 1e3, (double)T * N * 2 / best / 1e6);
}

int main() {
  const int T = 201216;
  uint16_t* d; cudaMalloc(&d, (size_t)T * 1024 * 2);
  for (int N : {1024, 768, 256}) {
    for (int th : {512, 256}) {
      run<0>(d, T, N, th, "8B/lane 4 rows x 64B (current)");
      run<1>(d, T, N, th, "16B/lane 8 rows x 64B");
      run<2>(d, T, N, th, "16B/lane 4 rows x 128B");
      run<3>(d, T, N, th, "16B/lane 2 rows x 256B");
      run<4>(d, T, N, th, "16B/lane 1 row x 512B");
      run<5>(d, T, N, th, "native: lane = row, 32 rows x 16B");
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
