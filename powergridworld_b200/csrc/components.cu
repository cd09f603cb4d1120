// K1-K5 + the lagged part of K7, fused: one thread per (env, agent) steps all of the
// agent's components in order, writes the agent's observation rows, its real power
// (the input of the power flow) and its reward.  HBM-bound streaming over
// structure-of-arrays state (env index fastest => fully coalesced 8-byte accesses).
//
// The event row -- every quantity that is shared by all envs at this step: PV profile
// value, building weather row, EV window lists, base feeder load, done flag -- is
// staged once per CTA into shared memory with a TMA bulk copy (cp.async.bulk ->
// UBLKCP) and then read as warp-uniform broadcasts.
//
// Replaces HOT LOOP 1-3 of gridworld/multiagent_env.py:165-181 and gridworld/base.py:125-137.
// Compiled with -fmad=false (see component_math.cuh).
#include "component_math.cuh"
#include "internal.cuh"

namespace pgw {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Stage `bytes` (multiple of 16) from global to shared through the TMA unit.
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                             uint64_t* mbar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(mbar))
      : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(mbar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(mbar)),
      "r"(parity)
      : "memory");
}

__global__ void __launch_bounds__(128) component_kernel(const CompParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t mbar;
  double* drow = reinterpret_cast<double*>(smem_raw);
  int32_t* irow = reinterpret_cast<int32_t*>(smem_raw + (size_t)p.dstride * sizeof(double));

  const int event = p.event_mode == 0 ? 0 : (*p.clock + 1);
  if (threadIdx.x == 0) {
    mbar_init(&mbar, 1);
    const uint32_t dbytes = (uint32_t)p.dstride * 8u, ibytes = (uint32_t)p.istride * 4u;
    mbar_expect_tx(&mbar, dbytes + ibytes);
    tma_bulk_g2s(drow, p.dtab + (size_t)event * p.dstride, dbytes, &mbar);
    if (ibytes) tma_bulk_g2s(irow, p.itab + (size_t)event * p.istride, ibytes, &mbar);
  }
  __syncthreads();
  mbar_wait(&mbar, 0);

  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int a = blockIdx.y;
  if (e < p.E) {
    AgentIO io;
    io.E = p.E;
    io.actions = p.actions;
    io.obs = p.obs;
    io.sd = p.sd;
    io.si = p.si;
    io.init_soc = p.init_soc;
    io.vmin = p.vmin;
    io.vmax = p.vmax;
    io.vbus = p.vbus;
    io.dpar = p.dpar;
    io.ipar = p.ipar;
    io.drow = drow;
    io.irow = irow;
    const pgw_agent ag = p.agents[a];
    const size_t ae = (size_t)a * p.E + e;
    if (p.event_mode == 0) {
      agent_reset(ag, p.comps, io, e);
      p.agent_p[ae] = 0.0;
      p.ep_ret[ae] = 0.0;
    } else {
      double pw, rw;
      agent_step(ag, p.comps, io, e, pw, rw);
      p.agent_p[ae] = pw;
      p.rew[ae] = rw;
      if (p.advance_clock) p.ep_ret[ae] += rw;    // otherwise the power-flow epilogue owns it
      if (a == 0) p.done[e] = drow[0] != 0.0 ? 1 : 0;
    }
  }

  if (p.advance_clock) {                          // last CTA of the step publishes the new clock
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int t = atomicAdd(p.ticket, 1u);
      if (t == gridDim.x * gridDim.y - 1) {
        *p.ticket = 0u;
        *p.clock = event;
        __threadfence();
      }
    }
  }
}

cudaError_t launch_components(const CompParams& p, int smem_bytes, cudaStream_t s) {
  const int threads = 128;
  dim3 grid((p.E + threads - 1) / threads, p.A);
  component_kernel<<<grid, threads, smem_bytes, s>>>(p);
  return cudaGetLastError();
}

}  // namespace pgw
