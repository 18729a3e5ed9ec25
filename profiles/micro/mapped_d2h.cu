// Microbenchmark for the e2e path: how fast can a KERNEL write the step's outputs
// into mapped pinned host memory, compared with one cudaMemcpyAsync D2H of the dense block?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mapped_d2h mapped_d2h.cu && ./mapped_d2h
// rows = agent rows (49152), K = 8 feature rows of 24 bytes per agent, cnt ~ uniform 1..8 (mean 4.4).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// dense copy, 16-byte pieces
__global__ void copy_dense(const float4* __restrict__ src, float4* __restrict__ dst, size_t n16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// only the valid rows of every agent (8-byte pieces: rows are 24 bytes), warp per 4 agents: lane -> (agent, piece)
__global__ void copy_valid(const float2* __restrict__ src, float2* __restrict__ dst, const int* __restrict__ cnt, int rows) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int a0 = w * 4; a0 < rows; a0 += nw * 4) {
    // 4 agents x 24 pieces (8 rows x 3 pieces) = 96 pieces, 3 per lane
    for (int q = lane; q < 96; q += 32) {
      const int a = a0 + q / 24, pc = q % 24;
      if (a < rows && pc < cnt[a] * 3) dst[(size_t)a * 24 + pc] = src[(size_t)a * 24 + pc];
    }
  }
}
int main() {
  const int rows = 49152, K = 8;
  const size_t bytes = (size_t)rows * K * 24;
  float* d_src; int* d_cnt; float *h_pin, *h_map, *d_map;
  CK(cudaMalloc(&d_src, bytes)); CK(cudaMalloc(&d_cnt, rows * 4));
  CK(cudaMemset(d_src, 1, bytes));
  std::vector<int> cnt(rows); size_t valid = 0;
  for (int i = 0; i < rows; i++) { cnt[i] = 1 + (rand() % 8); valid += cnt[i]; }
  CK(cudaMemcpy(d_cnt, cnt.data(), rows * 4, cudaMemcpyHostToDevice));
  CK(cudaHostAlloc(&h_pin, bytes, cudaHostAllocDefault));
  CK(cudaHostAlloc(&h_map, bytes, cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer(&d_map, h_map, 0));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, size_t moved, auto fn) {
    for (int i = 0; i < 3; i++) fn();
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    const int reps = 20;
    for (int i = 0; i < reps; i++) fn();
    CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-44s %8.1f us  %6.1f GB/s (%.2f MB)\n", name, ms * 1e3 / reps, moved / (ms / reps * 1e-3) / 1e9, moved / 1e6);
  };
  timeit("cudaMemcpyAsync D2H dense", bytes, [&] { CK(cudaMemcpyAsync(h_pin, d_src, bytes, cudaMemcpyDeviceToHost, st)); });
  for (int g : {64, 148, 296, 592, 1184})
    timeit((std::string("kernel dense -> mapped, grid ") + std::to_string(g)).c_str(), bytes,
           [&] { copy_dense<<<g, 256, 0, st>>>((const float4*)d_src, (float4*)d_map, bytes / 16); });
  for (int g : {148, 296, 592})
    timeit((std::string("kernel valid rows -> mapped, grid ") + std::to_string(g)).c_str(), valid * 24,
           [&] { copy_valid<<<g, 256, 0, st>>>((const float2*)d_src, (float2*)d_map, d_cnt, rows); });
  printf("valid fraction %.3f\n", (double)valid / (rows * K));
  return 0;
}
