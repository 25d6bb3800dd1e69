// Exclusive scan of `total` 32-bit counters in place, one 1024-thread block walking the array.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace asuna {

static __global__ void __launch_bounds__(1024) k_scan_exclusive(uint32_t* __restrict__ hist, uint32_t total) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < total; base += 1024) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < total ? hist[i] : 0u;
    uint32_t s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) warp_sums[warp] = s;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sums[lane], ws = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, ws, o);
        if (lane >= o) ws += t;
      }
      warp_sums[lane] = ws - w;  // exclusive
    }
    __syncthreads();
    uint32_t c = carry;
    uint32_t excl = c + warp_sums[warp] + s - v;
    if (i < total) hist[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
}

}  // namespace asuna
