// Second microbenchmark for the e2e export (see mapped_d2h.cu): the library's export kernels in isolation.
//   rows = 49152 agents, K = 8, 24-byte rows, cnt distribution like navigation-3 (mean ~2.9 of 8).
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void export_rows_rnd(const int* __restrict__ cnt, int* __restrict__ prev, const unsigned char* __restrict__ feat,
                                unsigned char* __restrict__ h_feat, long long rows, int row_bytes, int K, int RND) {
  // 8-byte pieces; the valid prefix is rounded UP to a multiple of RND bytes (the extra bytes are the zeros the
  // dense tensor holds there anyway): whole 32- / 64- / 128-byte lines instead of partial ones
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const int ab = K * row_bytes, ppa = ab / 8;
  for (long long a0 = w * 4; a0 < rows; a0 += nw * 4) {
    for (int q = lane; q < 4 * ppa; q += 32) {
      const long long a = a0 + q / ppa;
      if (a >= rows) continue;
      const int o = (q % ppa) * 8;
      const size_t g = (size_t)a * ab + o;
      int vb = cnt[a] * row_bytes, pb = prev[a] * row_bytes;
      vb = vb > pb ? vb : pb;
      // round the end up to RND relative to the ABSOLUTE address (line boundaries)
      const size_t end = ((size_t)a * ab + vb + RND - 1) / RND * RND;
      if (g < end && o < ab) *(uint2*)(h_feat + g) = *(const uint2*)(feat + g);
    }
    __syncwarp();
    if (lane < 4 && a0 + lane < rows) prev[a0 + lane] = cnt[a0 + lane];
  }
}
// LPA lanes per agent, each lane walks its agent's valid prefix in 8-byte pieces with stride LPA
template <int LPA>
__global__ void export_rows_lpa(const int* __restrict__ cnt, int* __restrict__ prev, const unsigned char* __restrict__ feat,
                                unsigned char* __restrict__ h_feat, long long rows, int row_bytes, int K) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
  const int ab = K * row_bytes;
  for (long long a = t / LPA; a < rows; a += nt / LPA) {
    const int sub = (int)(t % LPA);
    const int c = cnt[a], pv = prev[a];
    const int vb = (c > pv ? c : pv) * row_bytes;
    for (int o = sub * 8; o < vb; o += LPA * 8) {
      const size_t g = (size_t)a * ab + o;
      *(uint2*)(h_feat + g) = *(const uint2*)(feat + g);
    }
    if (sub == 0) prev[a] = c;
  }
}
template <int PB>
__global__ void export_rows(const int* __restrict__ cnt, int* __restrict__ prev, const unsigned char* __restrict__ feat,
                            unsigned char* __restrict__ h_feat, long long rows, int row_bytes, int K, int APW) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const int ab = K * row_bytes, ppa = ab / PB;
  for (long long a0 = w * APW; a0 < rows; a0 += nw * APW) {
    for (int q = lane; q < APW * ppa; q += 32) {
      const long long a = a0 + q / ppa;
      if (a >= rows) continue;
      const int o = (q % ppa) * PB;
      const int vb = cnt[a] * row_bytes, pb = prev[a] * row_bytes;
      const size_t g = (size_t)a * ab + o;
      if (o + PB <= vb) {
        if (PB == 16) *(uint4*)(h_feat + g) = *(const uint4*)(feat + g);
        else *(uint2*)(h_feat + g) = *(const uint2*)(feat + g);
      } else if (o < vb || o < pb) {
        for (int h8 = 0; h8 < PB; h8 += 8) {
          if (o + h8 < vb) *(uint2*)(h_feat + g + h8) = *(const uint2*)(feat + g + h8);
          else if (o + h8 < pb) *(uint2*)(h_feat + g + h8) = make_uint2(0u, 0u);
        }
      }
    }
    __syncwarp();
    if (lane < APW && a0 + lane < rows) prev[a0 + lane] = cnt[a0 + lane];
  }
}
__global__ void copy_dense(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
int main() {
  const int rows = 49152, K = 8;
  const size_t bytes = (size_t)rows * K * 24, small = 3800000;
  unsigned char *d_src, *h_map, *d_map, *d_small; int *d_cnt, *d_prev, *d_cnt2;
  CK(cudaMalloc(&d_src, bytes)); CK(cudaMalloc(&d_cnt, rows * 4)); CK(cudaMalloc(&d_cnt2, rows * 4)); CK(cudaMalloc(&d_prev, rows * 4));
  CK(cudaMalloc(&d_small, small + 64));
  CK(cudaMemset(d_src, 1, bytes));
  std::vector<int> c1(rows), c2(rows); size_t valid = 0;
  for (int i = 0; i < rows; i++) { c1[i] = 1 + (rand() % 5); c2[i] = std::max(1, std::min(8, c1[i] + (rand() % 3) - 1)); valid += c1[i]; }
  CK(cudaMemcpy(d_cnt, c1.data(), rows * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_cnt2, c2.data(), rows * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_prev, 0, rows * 4));
  CK(cudaHostAlloc(&h_map, bytes + small + 64, cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer(&d_map, h_map, 0));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const std::string& name, size_t moved, auto fn) {
    for (int i = 0; i < 4; i++) fn(i);
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    const int reps = 20;
    for (int i = 0; i < reps; i++) fn(i);
    CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-58s %8.1f us  %6.1f GB/s (%.2f MB)\n", name.c_str(), ms * 1e3 / reps, moved / (ms / reps * 1e-3) / 1e9, moved / 1e6);
  };
  for (int g : {148, 296, 1184})
    timeit("rows PB=8  APW=4 grid " + std::to_string(g), valid * 24, [&](int i) { export_rows<8><<<g, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K, 4); });
  for (int g : {148, 296, 1184})
    timeit("rows PB=16 APW=8 grid " + std::to_string(g), valid * 24, [&](int i) { export_rows<16><<<g, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K, 8); });
  timeit("rows PB=16 APW=4 grid 296", valid * 24, [&](int i) { export_rows<16><<<296, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K, 4); });
  timeit("rows LPA=1", valid * 24, [&](int i) { export_rows_lpa<1><<<296, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K); });
  timeit("rows LPA=2", valid * 24, [&](int i) { export_rows_lpa<2><<<296, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K); });
  timeit("rows LPA=4", valid * 24, [&](int i) { export_rows_lpa<4><<<296, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K); });
  timeit("rows LPA=8", valid * 24, [&](int i) { export_rows_lpa<8><<<296, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K); });
  timeit("rows LPA=8 grid 1184", valid * 24, [&](int i) { export_rows_lpa<8><<<1184, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K); });
  timeit("rows LPA=4 grid 74", valid * 24, [&](int i) { export_rows_lpa<4><<<74, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K); });
  for (int rnd : {8, 32, 64, 128})
    timeit("rows rounded up to " + std::to_string(rnd) + " B lines", valid * 24, [&](int i) { export_rows_rnd<<<296, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K, rnd); });
  for (int g : {74, 148, 592})
    timeit("dense small 3.8 MB kernel grid " + std::to_string(g), small, [&](int) { copy_dense<<<g, 256, 0, st>>>((const uint4*)d_small, (uint4*)(d_map + bytes), small / 16); });
  timeit("dense small 3.8 MB cudaMemcpyAsync", small, [&](int) { CK(cudaMemcpyAsync(h_map + bytes, d_small, small, cudaMemcpyDeviceToHost, st)); });
  timeit("rows PB=8 + dense kernel", valid * 24 + small, [&](int i) { export_rows<8><<<296, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K, 4); copy_dense<<<148, 256, 0, st>>>((const uint4*)d_small, (uint4*)(d_map + bytes), small / 16); });
  timeit("rows PB=8 + dense memcpy", valid * 24 + small, [&](int i) { export_rows<8><<<296, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K, 4); CK(cudaMemcpyAsync(h_map + bytes, d_small, small, cudaMemcpyDeviceToHost, st)); });
  cudaStream_t st2; CK(cudaStreamCreate(&st2));
  cudaEvent_t ev, ev2; CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ev2, cudaEventDisableTiming));
  timeit("rows PB=8 (stream A) || dense memcpy (stream B)", valid * 24 + small, [&](int i) {
    CK(cudaEventRecord(ev, st)); CK(cudaStreamWaitEvent(st2, ev, 0));
    export_rows<8><<<296, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K, 4);
    CK(cudaMemcpyAsync(h_map + bytes, d_small, small, cudaMemcpyDeviceToHost, st2));
    CK(cudaEventRecord(ev2, st2)); CK(cudaStreamWaitEvent(st, ev2, 0)); });
  timeit("rows PB=8 (stream A) || dense kernel (stream B)", valid * 24 + small, [&](int i) {
    CK(cudaEventRecord(ev, st)); CK(cudaStreamWaitEvent(st2, ev, 0));
    export_rows<8><<<296, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K, 4);
    copy_dense<<<148, 256, 0, st2>>>((const uint4*)d_small, (uint4*)(d_map + bytes), small / 16);
    CK(cudaEventRecord(ev2, st2)); CK(cudaStreamWaitEvent(st, ev2, 0)); });
  timeit("ONE kernel: rows PB=8 blocks + dense blocks", valid * 24 + small, [&](int i) {
    CK(cudaEventRecord(ev, st)); CK(cudaStreamWaitEvent(st2, ev, 0));
    export_rows<8><<<148, 256, 0, st>>>(i & 1 ? d_cnt2 : d_cnt, d_prev, d_src, d_map, rows, 24, K, 4);
    copy_dense<<<32, 256, 0, st2>>>((const uint4*)d_small, (uint4*)(d_map + bytes), small / 16);
    CK(cudaEventRecord(ev2, st2)); CK(cudaStreamWaitEvent(st, ev2, 0)); });
  printf("valid fraction %.3f\n", (double)valid / (rows * K));
  return 0;
}
