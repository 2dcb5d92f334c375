// norm_kernels.cuh -- BertNormalizer on the device, as a byte-rewriting pre-pass in front of the scan kernels.
//
// Replaces (paths relative to /root/reference/tokenizers/src):
//   normalizers/bert.rs:92-136          clean_text, handle_chinese_chars, strip_accents (NFD + drop Mn), lowercase
//   tokenizer/normalizer.rs:317-428     NormalizedString::transform: the alignment of every normalized character with
//                                       the original character it came from, which is what token offsets are made of
//
// Every step of BertNormalizer maps ONE character to a (possibly empty) sequence of characters without looking at its
// neighbours, so the whole normalizer is a table: code point -> UTF-8 bytes of its image (host_tables.cu composes it
// from the probed per-character facts of bert_tables.inc for the four flags).  The one context-dependent part of NFD,
// canonical reordering of combining marks, only permutes characters that strip_accents then drops: all 809 characters with
// a non-zero combining class but 83 count as Mn for the reference (probed, tools/gen_bert_tables.py).  One of those 83 right
// behind another combining character is the only place where order could matter: such a batch is refused (ERR_NORM_UNSUPPORTED
// -> B2T_ERR_UNSUPPORTED), never normalized differently.
//
//   N1 norm_count   per 2 KB page of the ORIGINAL batch: bytes of its image, characters it holds
//   (exclusive scans of both, one host read of the total)
//   N2 norm_write   the normalized batch, and for every normalized byte the index of its ORIGINAL character (src_char);
//                   document offsets in the normalized batch + the original character index of every document start
//   ... the ordinary pipeline runs on the normalized batch with byte offsets ...
//   N3 norm_offsets token (byte_start, byte_end) in the normalized document -> (char_start, char_end) in the ORIGINAL
//                   one: [src_char(first byte), src_char(last byte) + 1) -- the union of the alignments of the token's
//                   first and last character, exactly what the reference reports (pre_tokenizer.rs:198-263 + normalizer.rs
//                   alignments; probed: a dropped accent or control character inside a token widens it, one behind it does not)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "b2t_tables.h"
#include "pretok_logic.cuh"

namespace b2t {

enum { ERR_NORM_UNSUPPORTED = 16u };
constexpr int NORM_THREADS = 256;
constexpr int NORM_PER_THREAD = PAGE / NORM_THREADS;   // 8 bytes

struct NormTables {
  const uint16_t* blk;      // [0x110000 >> 7]: block of the code point
  const uint32_t* ent;      // [blocks][128]: kind (bits 0-1) | byte length (bits 2-7) | pool offset (bits 8-31)
  const uint8_t* pool;      // UTF-8 images
  const uint8_t* ascii;     // [128]: image of an ASCII character (always one ASCII character), 0xFF = dropped
};

__device__ __forceinline__ uint32_t norm_entry(const NormTables& T, uint32_t cp) {
  return __ldg(T.ent + (uint32_t)__ldg(T.blk + (cp >> 7)) * 128u + (cp & 127u));
}

// the character that starts at p (a lead byte): code point and byte length (bytes past the end read as 0)
__device__ __forceinline__ uint32_t norm_decode(const uint8_t* __restrict__ b, int64_t p, int64_t n, int* len) {
  const uint32_t b0 = __ldg(b + p);
  if (b0 < 0x80u) { *len = 1; return b0; }
  const int l = b0 < 0xE0u ? 2 : (b0 < 0xF0u ? 3 : 4);
  uint32_t cp = b0 & (0x7Fu >> l);
  for (int k = 1; k < l; ++k) cp = (cp << 6) | (p + k < n ? (__ldg(b + p + k) & 0x3Fu) : 0u);
  *len = l;
  return cp < 0x110000u ? cp : 0x10FFFFu;
}

// image length of the character at p (0 for continuation bytes); *is_lead tells whether p starts a character
__device__ __forceinline__ int norm_out_len(const NormTables& T, const uint8_t* __restrict__ b, int64_t p, int64_t n, bool* is_lead, uint32_t* err) {
  const uint32_t b0 = __ldg(b + p);
  if ((b0 & 0xC0u) == 0x80u) { *is_lead = false; return 0; }
  *is_lead = true;
  if (b0 < 0x80u) return __ldg(T.ascii + b0) != 0xFFu ? 1 : 0;
  int l;
  const uint32_t cp = norm_decode(b, p, n, &l);
  const uint32_t e = norm_entry(T, cp);
  switch (e & 3u) {
    case NORM_IDENT: return l;
    case NORM_REMOVE: return 0;
    case NORM_STRING: return (int)((e >> 2) & 63u);
    default: {
      // a combining character that survives strip_accents: NFD would have to order it against the combining character in
      // front of it, if there is one (dropped or not) -- that case is refused, everything else is the identity
      bool bad = (e & NORM_CCC_FLAG) != 0u;
      if (p > 0) {
        int64_t q = p - 1;
        while (q > 0 && (__ldg(b + q) & 0xC0u) == 0x80u && p - q < 4) --q;
        int l2;
        const uint32_t e2 = norm_entry(T, norm_decode(b, q, n, &l2));
        bad = bad || (e2 & 3u) == NORM_SURVIVOR || ((e2 & 3u) == NORM_REMOVE && (e2 & NORM_CCC_FLAG));
      }
      if (bad) atomicOr(err, ERR_NORM_UNSUPPORTED);
      return l;
    }
  }
}

// block-wide exclusive scan of one int per thread (256 threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int norm_block_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) { const int o = __shfl_up_sync(0xFFFFFFFFu, inc, s); if (lane >= s) inc += o; }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < NORM_THREADS / 32; ++w) { const int x = s_warp[w]; if (w < warp) base += x; tot += x; }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

// ------------------------------------------------------------------------------------------------ N1
__global__ void __launch_bounds__(NORM_THREADS) norm_count_kernel(const uint8_t* __restrict__ bytes, int64_t n, const NormTables T,
                                                                  uint32_t* __restrict__ page_out, uint32_t* __restrict__ page_chars, uint32_t* __restrict__ err) {
  __shared__ int s_warp[NORM_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * PAGE + (int64_t)threadIdx.x * NORM_PER_THREAD;
  int out = 0, chars = 0;
#pragma unroll
  for (int i = 0; i < NORM_PER_THREAD; ++i) {
    const int64_t p = base + i;
    if (p < n) { bool lead; out += norm_out_len(T, bytes, p, n, &lead, err); chars += lead ? 1 : 0; }
  }
  int tot_out, tot_chars;
  norm_block_scan(out, s_warp, &tot_out);
  norm_block_scan(chars, s_warp, &tot_chars);
  if (threadIdx.x == 0) { page_out[blockIdx.x] = (uint32_t)tot_out; page_chars[blockIdx.x] = (uint32_t)tot_chars; }
}

// ------------------------------------------------------------------------------------------------ N2
// page_out_base / page_char_base: exclusive scans of N1's counts (two-level: local + block, as the token-count scan).
__global__ void __launch_bounds__(NORM_THREADS) norm_write_kernel(const uint8_t* __restrict__ bytes, int64_t n, const NormTables T,
                                                                  const unsigned long long* __restrict__ out_lexcl, const unsigned long long* __restrict__ out_bexcl,
                                                                  const unsigned long long* __restrict__ chr_lexcl, const unsigned long long* __restrict__ chr_bexcl, int scan_block,
                                                                  const uint32_t* __restrict__ doc_bits, const uint64_t* __restrict__ doc_off, uint32_t n_docs,
                                                                  uint8_t* __restrict__ out, uint32_t* __restrict__ src_char,
                                                                  uint64_t* __restrict__ doc_off_out, uint32_t* __restrict__ doc_char0, uint32_t* __restrict__ err) {
  __shared__ int s_warp[NORM_THREADS / 32];
  const int64_t t = blockIdx.x;
  const int64_t base = t * PAGE + (int64_t)threadIdx.x * NORM_PER_THREAD;
  int lens[NORM_PER_THREAD];
  bool leads[NORM_PER_THREAD];
  int n_out = 0, chars = 0;
#pragma unroll
  for (int i = 0; i < NORM_PER_THREAD; ++i) {
    const int64_t p = base + i;
    lens[i] = 0; leads[i] = false;
    if (p < n) { lens[i] = norm_out_len(T, bytes, p, n, &leads[i], err); n_out += lens[i]; chars += leads[i] ? 1 : 0; }
  }
  int tot;
  const int out_excl = norm_block_scan(n_out, s_warp, &tot);
  const int chr_excl = norm_block_scan(chars, s_warp, &tot);
  unsigned long long o = out_lexcl[t] + out_bexcl[t / scan_block] + (unsigned long long)out_excl;
  unsigned long long c = chr_lexcl[t] + chr_bexcl[t / scan_block] + (unsigned long long)chr_excl;
  // the 8 bytes of this thread lie in one word pair of the document bitmap
  const uint32_t dsw = base < n + 1 ? __ldg(doc_bits + (base >> 5)) >> (base & 31) : 0u;
#pragma unroll
  for (int i = 0; i < NORM_PER_THREAD; ++i) {
    const int64_t p = base + i;
    if (p > n) break;
    if ((dsw >> i) & 1u) {
      // every document that starts at byte p (several if some are empty): doc_off is sorted, find the first by bisection
      uint32_t lo = 0, hi = n_docs;
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((int64_t)__ldg(doc_off + mid) < p) lo = mid + 1; else hi = mid; }
      for (uint32_t d = lo; d <= n_docs && (int64_t)__ldg(doc_off + d) == p; ++d) { doc_off_out[d] = o; doc_char0[d] = (uint32_t)c; }
    }
    if (p >= n || !leads[i]) continue;
    const uint32_t b0 = __ldg(bytes + p);
    if (b0 < 0x80u) {
      const uint32_t img = __ldg(T.ascii + b0);
      if (img != 0xFFu) { out[o] = (uint8_t)img; src_char[o] = (uint32_t)c; }
    } else {
      int l;
      const uint32_t cp = norm_decode(bytes, p, n, &l);
      const uint32_t e = norm_entry(T, cp);
      const uint32_t kind = e & 3u;
      if (kind == NORM_STRING) {
        const uint8_t* __restrict__ src = T.pool + (e >> 8);
        for (int k = 0; k < lens[i]; ++k) { out[o + k] = __ldg(src + k); src_char[o + k] = (uint32_t)c; }
      } else if (kind != NORM_REMOVE) {
        for (int k = 0; k < l; ++k) { out[o + k] = p + k < n ? __ldg(bytes + p + k) : 0; src_char[o + k] = (uint32_t)c; }
      }
    }
    o += (unsigned long long)lens[i];
    c += 1ull;
  }
}

// ------------------------------------------------------------------------------------------------ N3
// One warp per document: byte offsets in the normalized document -> character offsets in the original one.
__global__ void norm_offsets_kernel(const uint64_t* __restrict__ row_ptr, uint32_t n_docs, unsigned long long token_base,
                                    const uint64_t* __restrict__ doc_off_norm, const uint32_t* __restrict__ doc_char0,
                                    const uint32_t* __restrict__ src_char, uint2* __restrict__ offsets) {
  const uint32_t d = (uint32_t)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (d >= n_docs) return;
  const uint64_t a = row_ptr[d] - token_base, b = row_ptr[d + 1] - token_base;
  const uint64_t base = doc_off_norm[d];
  const uint32_t c0 = doc_char0[d];
  for (uint64_t i = a + lane; i < b; i += 32) {
    const uint2 o = offsets[i];
    uint2 r = make_uint2(0u, 0u);
    if (o.y > o.x) r = make_uint2(__ldg(src_char + base + o.x) - c0, __ldg(src_char + base + o.y - 1) + 1u - c0);
    offsets[i] = r;
  }
}

}  // namespace b2t
