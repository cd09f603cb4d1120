// Microbenchmark: FP64 FMA throughput of one SM (independent chains), DMMA m8n8k4 throughput,
// and broadcast LDS.128 throughput.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#include <mma.h>

template <int CHAINS>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-9 + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma_kernel(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = threadIdx.x * 1e-9f + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = fmaf(acc[c], a, b);
  }
  float s = 0;
#pragma unroll
  for (int c = 0; c < 16; ++c) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
  using namespace nvcuda::wmma;
  fragment<matrix_a, 8, 8, 4, double, row_major> a;
  fragment<matrix_b, 8, 8, 4, double, col_major> b;
  fragment<accumulator, 8, 8, 4, double> c[4];
  fill_fragment(a, 1.0000001);
  fill_fragment(b, 0.9999999);
  for (int q = 0; q < 4; ++q) fill_fragment(c[q], 0.0);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 4; ++q) mma_sync(c[q], a, b, c[q]);
  }
  double s = 0;
  for (int q = 0; q < 4; ++q)
    for (int i = 0; i < c[q].num_elements; ++i) s += c[q].x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void lds_bcast_kernel(double* out, int iters) {
  __shared__ double2 tab[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = make_double2(i, -i);
  __syncthreads();
  double2 s = make_double2(0, 0);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      double2 v;
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y)
                   : "r"((unsigned)__cvta_generic_to_shared(&tab[(i + q * 7) & 255])));
      s.x += v.x; s.y += v.y;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s.x + s.y;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double* out;
  cudaMalloc(&out, 148 * 1024 * 8 * 4);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double clk = clk_khz * 1e3;
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    float ms = time_ms([&] { dfma_kernel<16><<<148, warps * 32>>>(out, iters, 1.0000001, 1e-9); });
    double fma_per_clk_sm = (double)warps * 32 * 16 * iters / (ms * 1e-3 * clk);
    printf("DFMA  warps/SM=%2d chains=16: %.3f ms  -> %.1f DFMA/clk/SM (%.1f TFLOP/s chip)\n", warps, ms,
           fma_per_clk_sm, fma_per_clk_sm * 2 * 148 * clk / 1e12);
  }
  {
    float ms = time_ms([&] { dfma_kernel<1><<<148, 32>>>(out, iters, 1.0000001, 1e-9); });
    printf("DFMA dependent-chain latency: %.1f clk\n", ms * 1e-3 * clk / iters);
  }
  {
    float ms = time_ms([&] { ffma_kernel<<<148, 1024>>>((float*)out, iters, 1.0000001f, 1e-9f); });
    double r = 1024.0 * 16 * iters / (ms * 1e-3 * clk);
    printf("FFMA  warps/SM=32: %.1f FFMA/clk/SM\n", r);
  }
  for (int warps : {4, 8, 16}) {
    float ms = time_ms([&] { dmma_kernel<<<148, warps * 32>>>(out, iters); });
    double mma = (double)warps * 4 * iters;   // per SM
    printf("DMMA m8n8k4 warps/SM=%2d: %.3f ms -> %.3f mma/clk/SM = %.1f FMA/clk/SM (%.1f TFLOP/s chip)\n", warps,
           ms, mma / (ms * 1e-3 * clk), mma * 256 / (ms * 1e-3 * clk), mma * 256 / (ms * 1e-3 * clk) * 2 * 148 * clk / 1e12);
  }
  for (int warps : {8}) {
    float ms = time_ms([&] { lds_bcast_kernel<<<148, warps * 32>>>(out, iters); });
    printf("LDS.128 broadcast warps/SM=%d: %.2f clk per warp-instruction per SM\n", warps,
           ms * 1e-3 * clk / ((double)warps * 16 * iters));
  }
  printf("clock %.0f MHz\n", clk / 1e6);
  return 0;
}
