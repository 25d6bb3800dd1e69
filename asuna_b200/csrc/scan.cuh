// Exclusive scan of `total` 32-bit counters in place: one 1024-thread block walks the array 8192 counters at a time
// (8 consecutive counters per thread, so a warp reads 1 KB contiguously; two block barriers per step).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace asuna {

static __global__ void __launch_bounds__(1024) k_scan_exclusive(uint32_t* __restrict__ hist, uint32_t total) {
  constexpr int kItems = 8;
  __shared__ uint32_t warp_sums[2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t carry = 0;  // every thread tracks the running total (same value in all of them)
  int buf = 0;
  for (uint32_t base = 0; base < total; base += 1024 * kItems, buf ^= 1) {
    const uint32_t i0 = base + threadIdx.x * kItems;
    uint32_t v[kItems], sum = 0;
#pragma unroll
    for (int k = 0; k < kItems; k++) {
      v[k] = i0 + k < total ? hist[i0 + k] : 0u;
      sum += v[k];
    }
    uint32_t s = sum;  // inclusive scan of the thread totals inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) warp_sums[buf][warp] = s;
    __syncthreads();
    const uint32_t w = warp_sums[buf][lane];  // every warp scans the 32 warp totals itself: no second barrier
    uint32_t ws = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, ws, o);
      if (lane >= o) ws += t;
    }
    const uint32_t block_total = __shfl_sync(0xFFFFFFFFu, ws, 31);
    const uint32_t warp_excl = __shfl_sync(0xFFFFFFFFu, ws - w, warp);
    uint32_t run = carry + warp_excl + s - sum;
#pragma unroll
    for (int k = 0; k < kItems; k++) {
      if (i0 + k < total) hist[i0 + k] = run;
      run += v[k];
    }
    carry += block_total;
    // warp_sums is double-buffered: the next step writes the other buffer, so one barrier per step is enough
  }
}

}  // namespace asuna
