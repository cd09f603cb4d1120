// What a kernel launch costs between two CUDA events on the launch stream, as a function of the
// things step_fused_kernel asks for: parameter bytes, dynamic shared memory, a full register file
// per CTA, a TMEM allocation, graph vs direct launch, and a different kernel (the L2 flush of
// bench.py) in front of it.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/microbench/launch_overhead.cu -o tools/_build/launch_overhead
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

struct Small { int x[16]; };
struct Big { int x[2000]; };                           // ~8 kB like FusedParams + Tc2Consts + Tc2Polish

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename P, bool TMEM, int REGS_HOG>
__global__ void __launch_bounds__(512, 1) k(const __grid_constant__ P p, int* out, long long* stamps) {
  extern __shared__ unsigned char sm[];
  __shared__ uint32_t tbase;
  if (stamps && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    stamps[blockIdx.x * 2] = (long long)t;
  }
  if (TMEM) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(64u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(64u) : "memory");
  }
  int acc = p.x[threadIdx.x & 15];
  if (REGS_HOG) {                                      // keep ~100 live registers
    int v[96];
#pragma unroll
    for (int i = 0; i < 96; ++i) asm volatile("mov.u32 %0, %1;" : "=r"(v[i]) : "r"(acc + i));
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 96; ++i) asm volatile("add.u32 %0, %0, %1;" : "+r"(acc) : "r"(v[i]));
  }
  if (acc == 0x7fffffff) out[0] = acc + sm[threadIdx.x];
  if (stamps && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    stamps[blockIdx.x * 2 + 1] = (long long)t;
  }
}

__global__ void fill(uint4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_uint4(1, 2, 3, 4);
}

template <typename F>
static void run(const char* name, F launch, cudaStream_t s, bool flush, uint4* fb, size_t fn, bool graph) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaGraphExec_t exec = nullptr;
  if (graph) {
    cudaGraph_t g;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    launch();
    cudaStreamEndCapture(s, &g);
    cudaGraphInstantiate(&exec, g, 0);
    cudaGraphDestroy(g);
  }
  std::vector<float> t;
  for (int i = 0; i < 120; ++i) {
    if (flush) fill<<<148 * 8, 256, 0, s>>>(fb, fn);
    cudaEventRecord(e0, s);
    if (graph) cudaGraphLaunch(exec, s); else launch();
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (i >= 20) t.push_back(ms * 1e3f);
  }
  std::sort(t.begin(), t.end());
  printf("%-58s flush=%d graph=%d  median %6.2f us  min %6.2f us\n", name, (int)flush, (int)graph, t[t.size() / 2], t[0]);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(err));
}

int main() {
  cudaStream_t s;
  cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  int* out;
  cudaMalloc(&out, 64);
  long long* stamps;
  cudaMalloc(&stamps, 148 * 16);
  size_t fn = (256u << 20) / 16;
  uint4* fb;
  cudaMalloc(&fb, fn * 16);
  Small sp{};
  Big bp{};
  const int big_smem = 120 * 1024;
  cudaFuncSetAttribute(k<Small, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
  cudaFuncSetAttribute(k<Big, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
  cudaFuncSetAttribute(k<Big, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
  cudaFuncSetAttribute(k<Big, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
  cudaFuncSetAttribute(k<Small, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big_smem);
  for (int flush = 0; flush < 2; ++flush)
    for (int graph = 0; graph < 2; ++graph) {
      run("empty stream (two events back to back)", [&] {}, s, flush, fb, fn, false);
      run("small params, 128 CTAs x 512, no smem", [&] { k<Small, false, 0><<<128, 512, 0, s>>>(sp, out, nullptr); }, s, flush, fb, fn, graph);
      run("small params, 120 kB smem", [&] { k<Small, false, 0><<<128, 512, big_smem, s>>>(sp, out, nullptr); }, s, flush, fb, fn, graph);
      run("8 kB params, no smem", [&] { k<Big, false, 0><<<128, 512, 0, s>>>(bp, out, nullptr); }, s, flush, fb, fn, graph);
      run("8 kB params, 120 kB smem", [&] { k<Big, false, 0><<<128, 512, big_smem, s>>>(bp, out, nullptr); }, s, flush, fb, fn, graph);
      run("8 kB params, 120 kB smem, TMEM alloc", [&] { k<Big, true, 0><<<128, 512, big_smem, s>>>(bp, out, nullptr); }, s, flush, fb, fn, graph);
      run("8 kB params, 120 kB smem, TMEM, ~128 regs", [&] { k<Big, true, 1><<<128, 512, big_smem, s>>>(bp, out, nullptr); }, s, flush, fb, fn, graph);
      run("small params, 120 kB smem, TMEM, ~128 regs", [&] { k<Small, true, 1><<<128, 512, big_smem, s>>>(sp, out, nullptr); }, s, flush, fb, fn, graph);
      run("same with entry/exit stamps", [&] { k<Small, true, 1><<<128, 512, big_smem, s>>>(sp, out, stamps); }, s, flush, fb, fn, graph);
      long long h[296];
      cudaMemcpy(h, stamps, sizeof(h), cudaMemcpyDeviceToHost);
      long long lo = h[0], hi = h[1];
      for (int i = 0; i < 128; ++i) { lo = std::min(lo, h[2 * i]); hi = std::max(hi, h[2 * i + 1]); }
      printf("    first entry -> last exit of that launch: %.2f us\n", (hi - lo) * 1e-3);
    }
  return 0;
}
