// added_emul.cpp -- TEST ONLY: runs the device code of tokenizers_b200/csrc/added_kernels.cuh (A1 candidate scan, A2
// per-document resolution) on the host, one "thread" after the other, so that the added-token extraction the kernels
// implement can be fuzzed against the host logic (tokenizers_b200/added.py, itself pinned to the wheel) without a GPU.
// The CUDA keywords and intrinsics the header uses are shimmed below; the header itself is compiled unchanged.
//   g++ -O2 -std=c++17 -I/usr/local/cuda/include -Wno-attributes -shared -fPIC -o libadded_emul.so added_emul.cpp
#include <stdint.h>
#include <string.h>
#include <cuda_runtime.h>

struct Dim3e { unsigned x = 0, y = 0, z = 0; };
static Dim3e blockIdx, threadIdx, blockDim, gridDim;
#define __syncthreads() ((void)0)
#define __launch_bounds__(...)
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p |= v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline unsigned atomicExch(unsigned* p, unsigned v) { unsigned o = *p; *p = v; return o; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31u; return s ? (lo >> s) | (hi << (32u - s)) : lo; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }   // (only the kernel wrapper uses it; not called here)
#undef __shared__
#define __shared__ static

#include "../../tokenizers_b200/csrc/added_kernels.cuh"

using namespace b2t;

// Runs A1 + A2 over a packed batch.  hard_bits must arrive holding the document-start bits (the engine copies doc_bits
// there), inner_bits / added_bits zeroed, head filled with 0xFFFFFFFF.  Returns the number of pool entries used.
extern "C" uint32_t b2t_emul_added(const uint8_t* bytes, uint64_t n, const uint64_t* doc_off, uint32_t n_docs,
                                   const uint8_t* tok_bytes, const uint32_t* tok_off, const uint32_t* tok_id, const uint8_t* tok_flags,
                                   const uint32_t* set_begin, const uint32_t* first_bits, const uint32_t* pair_bits, const uint32_t* cls_rust,
                                   uint32_t n_first, const uint32_t* first_bcast,
                                   uint32_t* cand0, uint32_t* cand1, uint32_t* cand_any,
                                   uint32_t* hard_bits, uint32_t* inner_bits, uint32_t* added_bits, uint32_t* head, uint32_t* pool /* uint2 */,
                                   uint32_t pool_cap, uint32_t* err) {
  AddedTables T;
  T.tok_bytes = tok_bytes; T.tok_off = tok_off; T.tok_id = tok_id; T.tok_flags = tok_flags;
  T.set_begin[0] = set_begin[0]; T.set_begin[1] = set_begin[1]; T.set_begin[2] = set_begin[2];
  T.first_bits = first_bits; T.pair_bits = pair_bits; T.cls_rust = cls_rust;
  T.n_first = n_first;
  for (int i = 0; i < 4; ++i) T.first_bcast[i] = first_bcast[i];
  // A1, chunk by chunk (added_scan_kernel minus its thread indexing and the warp vote)
  const int64_t n_chunks = (int64_t)n / CHUNK + 1;
  for (int64_t c = 0; c < n_chunks; ++c) {
    const int64_t base = c * CHUNK;
    uint32_t m0 = 0u, m1 = 0u;
    if (base < (int64_t)n) { added_scan_chunk(bytes, (int64_t)n, T, first_bits, c, base, m0, m1); cand0[c] = m0; cand1[c] = m1; }
    if (m0 | m1) cand_any[c >> 5] |= 1u << (c & 31);
  }
  // A2, document by document
  uint32_t used = 0;
  AddedOut o;
  o.hard_bits = hard_bits; o.inner_bits = inner_bits; o.added_bits = added_bits; o.head = head; o.pool = reinterpret_cast<uint2*>(pool);
  o.pool_used = &used; o.pool_cap = pool_cap; o.err = err;
  blockDim.x = 128;
  for (uint32_t d = 0; d < n_docs; ++d) {
    blockIdx.x = d / 128; threadIdx.x = d % 128;
    added_resolve_kernel(bytes, doc_off, n_docs, T, cand0, cand1, cand_any, o);
  }
  return used;
}
