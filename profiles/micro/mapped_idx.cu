// Microbenchmark: PCIe cost of writing a FRACTION of 49152 32-byte rows (nbr_idx rows) into mapped host memory,
// each as two 16-byte stores by two adjacent lanes, against the dense 1.57 MB block.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__global__ void rows(const uint4* __restrict__ src, uint4* __restrict__ dst, const unsigned char* __restrict__ on, int n) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < 2 * n; q += gridDim.x * blockDim.x)
    if (on[q >> 1]) dst[q] = src[q];
}
int main() {
  const int n = 49152;
  uint4 *d_src, *h_map, *d_map; unsigned char* d_on;
  CK(cudaMalloc(&d_src, n * 32)); CK(cudaMemset(d_src, 1, n * 32)); CK(cudaMalloc(&d_on, n));
  CK(cudaHostAlloc(&h_map, n * 32, cudaHostAllocMapped)); CK(cudaHostGetDevicePointer(&d_map, h_map, 0));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int pct : {100, 60, 40, 30, 20, 10, 5}) {
    std::vector<unsigned char> on(n);
    size_t cntv = 0;
    for (int i = 0; i < n; i++) { on[i] = (rand() % 100) < pct; cntv += on[i]; }
    CK(cudaMemcpy(d_on, on.data(), n, cudaMemcpyHostToDevice));
    for (int i = 0; i < 3; i++) rows<<<148, 256, 0, st>>>(d_src, d_map, d_on, n);
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < 20; i++) rows<<<148, 256, 0, st>>>(d_src, d_map, d_on, n);
    CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%3d %% of the rows: %7.1f us  (%.2f MB)\n", pct, ms * 1e3 / 20, cntv * 32 / 1e6);
  }
  return 0;
}
