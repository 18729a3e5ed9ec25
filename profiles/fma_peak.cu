// fp32 FMA issue-rate microbenchmark for the actor kernel's roofline denominator (sm_100a).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fma_peak fma_peak.cu && ./fma_peak
// Variants: scalar FFMA with register operands, scalar FFMA with a constant (uniform) operand,
// packed FFMA2 (fma.rn.f32x2) with register operands, FFMA2 with a constant operand pair.
#include <cuda_runtime.h>
#include <cstdio>

struct Consts { float2 c[64]; };

template <int MODE>
__global__ void __launch_bounds__(256) k(const __grid_constant__ Consts w, float* out, int iters, float seed) {
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
  float2 x = make_float2(seed * 0.5f, seed * 0.25f);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (MODE == 0) { a[i].x = fmaf(a[i].x, x.x, x.y); a[i].y = fmaf(a[i].y, x.y, x.x); }
        if (MODE == 1) { a[i].x = fmaf(a[i].x, w.c[r * 8 + i].x, x.y); a[i].y = fmaf(a[i].y, w.c[r * 8 + i].y, x.x); }
        if (MODE == 2) a[i] = __ffma2_rn(a[i], x, x);
        if (MODE == 3) a[i] = __ffma2_rn(a[i], w.c[r * 8 + i], x);
        if (MODE == 4) { const float ws = w.c[r * 8 + i].x; a[i] = __ffma2_rn(make_float2(ws, ws), a[i], x); }
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
  if (s == 12345.f) out[0] = s;
}

template <int MODE> void run(const char* name, const Consts& w, float* d) {
  const int iters = 2000, grid = 148 * 8, block = 256;
  k<MODE><<<grid, block>>>(w, d, 10, 1.f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, block>>>(w, d, iters, 1.f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)grid * block * iters * 64 * 2;      // scalar-equivalent FMAs
  printf("%-34s %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM @1.965GHz\n", name, ms, 2 * fma / ms / 1e9,
         fma / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
  Consts w; for (int i = 0; i < 64; i++) w.c[i] = make_float2(1.0f + i * 1e-7f, 1.0f - i * 1e-7f);
  float* d; cudaMalloc(&d, 4);
  run<0>("FFMA  reg operands", w, d);
  run<1>("FFMA  constant operand", w, d);
  run<2>("FFMA2 reg operands", w, d);
  run<3>("FFMA2 constant operand pair", w, d);
  run<4>("FFMA2 broadcast scalar constant", w, d);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
