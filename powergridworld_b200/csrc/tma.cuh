// TMA bulk-copy staging helpers (cp.async.bulk + mbarrier; SASS: UBLKCP / SYNCS).
#pragma once
#include <stdint.h>

namespace pgw {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Stage `bytes` (multiple of 16, 16-byte aligned on both sides) from global to shared.
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

// The last kernel of a step advances the episode clock on the device, so that a step has
// constant launch parameters (CUDA-graph replayable).  Every CTA reads the clock at entry and
// then takes a ticket (one thread); the CTA that took the LAST ticket knows that all the others
// have read the old value, and writes the new one when it is done.  The ticket is taken at
// entry -- its round trip to L2 hides behind the kernel's work, nothing waits for it until
// the very end -- and its increment carries a data dependence on the value read, so that the
// read is performed before the ticket is taken.
__device__ __forceinline__ unsigned int clock_take_ticket(unsigned int* ticket, int clock_read) {
  const unsigned int inc = 1u + ((unsigned int)clock_read >> 31);   // 1: the clock is never negative
  return atomicAdd(ticket, inc);
}
__device__ __forceinline__ void clock_advance_if_last(unsigned int my_ticket, unsigned int* ticket,
                                                      int* clock, int clock_read,
                                                      unsigned int num_ctas) {
  if (my_ticket == num_ctas - 1) {
    *ticket = 0u;
    *clock = clock_read + 1;
  }
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

}  // namespace pgw
