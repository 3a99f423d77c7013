// sm_100a PTX wrappers: mbarrier transactions and 1-D bulk (TMA) copies from global to shared memory.
#pragma once
#include <stdint.h>

namespace ipp {

namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0u;
}
// A waiting CONSUMER warp must not compete for issue slots with the warps that compute or — worse — with the one
// producer warp every other warp depends on: poll, then sleep between polls (try_wait alone, or try_wait with a
// suspend-time hint = NANOSLEEP.SYNCS, re-polls on every barrier event of the CTA: measured 19-41 % of all executed
// warp instructions of the map kernel).  The sleep bounds the wake-up delay of a starved warp to `ns`.
#ifndef IPP_TMA_SLEEP
#define IPP_TMA_SLEEP 256u
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t ns = IPP_TMA_SLEEP) {
  if (mbar_try_wait(bar, parity)) return;
  do {
    __nanosleep(ns);
  } while (!mbar_try_wait(bar, parity));
}
// The producer's own waits are on the critical path of every other warp: spin.
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// named barrier among `threads` threads (whole warps) of the CTA
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// pull `bytes` (multiple of 16) of global memory into L2 without a destination: the later plain loads of the same
// bytes then see L2 latency instead of DRAM latency
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
}  // namespace ptx

}  // namespace ipp
