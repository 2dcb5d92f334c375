// prefix_kernels.cuh -- ByteLevel(add_prefix_space = true): pre_tokenizers/byte_level.rs:121-125 prepends " " to every
// non-empty sequence that does not already start with one (normalizer.rs:503-514; the new byte is aligned to the
// first original character).  On the device the batch is re-packed with the space physically inserted, the normal
// pipeline runs on the re-packed buffer, and the emit step maps offsets back (model_kernels.cuh, `prefix_bits`).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2t {

constexpr int PFX_BLOCK = 1024;

__device__ __forceinline__ uint32_t pfx_flag(const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ doc_off, uint32_t d) {
  const uint64_t a = doc_off[d], b = doc_off[d + 1];
  return (b > a && bytes[a] != ' ') ? 1u : 0u;
}

// (1) per block of 1024 documents: exclusive scan of the "gets a prefix" flags + block total
__global__ void __launch_bounds__(PFX_BLOCK) pfx_scan_block_kernel(const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ doc_off,
                                                                  uint32_t n_docs, uint32_t* __restrict__ local_excl, uint32_t* __restrict__ block_tot) {
  __shared__ uint32_t s_w[32];
  const uint32_t d = blockIdx.x * PFX_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t f = d < n_docs ? pfx_flag(bytes, doc_off, d) : 0u;
  uint32_t inc = f;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) { uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, s); if (lane >= s) inc += o; }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_w[lane], wi = w;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { uint32_t o = __shfl_up_sync(0xFFFFFFFFu, wi, s); if (lane >= s) wi += o; }
    s_w[lane] = wi - w;  // exclusive over warps
    if (lane == 31) block_tot[blockIdx.x] = wi;
  }
  __syncthreads();
  if (d < n_docs) local_excl[d] = s_w[warp] + inc - f;
}

// (2) one block: exclusive scan of the block totals (in place) + grand total
__global__ void __launch_bounds__(PFX_BLOCK) pfx_scan_top_kernel(uint32_t* __restrict__ block_tot, uint32_t n_blocks, unsigned long long* __restrict__ total) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_run;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < n_blocks; base += PFX_BLOCK) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n_blocks ? block_tot[i] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, s); if (lane >= s) inc += o; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_w[lane], wi = w;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) { uint32_t o = __shfl_up_sync(0xFFFFFFFFu, wi, s); if (lane >= s) wi += o; }
      s_w[lane] = wi - w;
    }
    __syncthreads();
    const uint32_t run = s_run;
    if (i < n_blocks) block_tot[i] = run + s_w[warp] + inc - v;
    __syncthreads();
    if (threadIdx.x == PFX_BLOCK - 1) s_run = run + s_w[warp] + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_run;
}

// (3) new document offsets, prefix bitmap
__global__ void pfx_offsets_kernel(const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ doc_off, uint32_t n_docs,
                                   const uint32_t* __restrict__ local_excl, const uint32_t* __restrict__ block_excl,
                                   const unsigned long long* __restrict__ total, uint64_t* __restrict__ new_off,
                                   uint32_t* __restrict__ prefix_bits) {
  const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d > n_docs) return;
  if (d == n_docs) { new_off[d] = doc_off[d] + *total; return; }
  const uint64_t shift = (uint64_t)local_excl[d] + block_excl[d / PFX_BLOCK];
  const uint64_t p = doc_off[d] + shift;
  new_off[d] = p;
  if (pfx_flag(bytes, doc_off, d)) atomicOr(&prefix_bits[p >> 5], 1u << (p & 31));
}

// (4) copy: one warp per document
__global__ void pfx_copy_kernel(const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ doc_off, const uint64_t* __restrict__ new_off,
                                uint32_t n_docs, uint8_t* __restrict__ out) {
  const uint32_t d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (d >= n_docs) return;
  const uint64_t a = doc_off[d], len = doc_off[d + 1] - a;
  uint64_t q = new_off[d];
  const bool pre = (new_off[d + 1] - q) != len;
  if (pre) { if (lane == 0) out[q] = ' '; ++q; }
  for (uint64_t i = lane; i < len; i += 32) out[q + i] = bytes[a + i];
}

}  // namespace b2t
