// pretok_kernels.cuh -- K0 doc_mark, K1 pretok_scan (the HBM-roofline-graded kernel), K1b page_scan.
//
// Data layout in HBM (one batch = `n` packed UTF-8 bytes, documents back to back):
//   bytes      u8 [n]                input, read once by K1 with 16-byte loads
//   doc_bits   u32[n/32 + 2]         bit p set <=> some document starts at byte p (K0; bit n is the end sentinel)
//   start_bits u32[n/32 + 2]         bit p set <=> a pre-token (split) starts at byte p            (K1 output)
//   drop_bits  u32[n/32 + 2]         Whitespace only: the split starting at p is removed whitespace (K1 output)
//   page_sum   u64[n/2048 + 1]       per 2 KB page: #chars, #kept splits, and the same counted from the last doc start
//   page_carry u64[n/2048 + 1]       K1b: #chars (low 32) / #kept splits (high 32) of the document that spans into the
//                                    page, counted from that document's start to the page start
//   page_first_doc u32[n/2048 + 1]   K0: first document d with doc_off[d] >= page start
// K1 algorithmic traffic: n bytes in + n/8 (doc_bits) in + n/8 (start_bits) out = 1.25 B per input byte.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pretok_logic.cuh"

namespace b2t {

struct ByteAtGlobal {
  const uint8_t* __restrict__ p;
  int64_t n;
  __device__ __forceinline__ uint32_t operator()(int64_t i) const { return (i >= 0 && i < n) ? (uint32_t)__ldg(p + i) : 0u; }
};

// ------------------------------------------------------------------------------------------------ K0
// One thread per document boundary d in [0, n_docs]: marks the doc start bit and fills page_first_doc.
__global__ void doc_mark_kernel(const uint64_t* __restrict__ doc_off, uint32_t n_docs, uint32_t* __restrict__ doc_bits,
                                uint32_t* __restrict__ page_first_doc) {
  uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d > n_docs) return;
  uint64_t pos = doc_off[d];
  atomicOr(&doc_bits[pos >> 5], 1u << (pos & 31));
  uint64_t p_hi = pos / PAGE;
  uint64_t p_lo = (d == 0) ? 0 : doc_off[d - 1] / PAGE + 1;
  for (uint64_t p = p_lo; p <= p_hi; ++p) page_first_doc[p] = d;
}

// ------------------------------------------------------------------------------------------------ K1
__device__ __forceinline__ void load_chunk_words(const uint8_t* __restrict__ bytes, int64_t base, int64_t n, uint32_t w[8]) {
  if (base + CHUNK <= n) {
    const uint4* q = reinterpret_cast<const uint4*>(bytes + base);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t x = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int64_t p = base + j * 4 + k;
        if (p < n) x |= (uint32_t)__ldg(bytes + p) << (8 * k);
      }
      w[j] = x;
    }
  }
}

template <int KIND>
__device__ __forceinline__ ChunkMasks classify_global(const uint8_t* __restrict__ bytes, int64_t n, int64_t chunk,
                                                      const uint32_t* __restrict__ cls_tbl) {
  ChunkMasks m;
  int64_t base = chunk * CHUNK;
  if (chunk < 0 || base >= n) {
    m.lead = m.L = m.N = m.S = m.SP = m.NL = m.AP = 0u;
    return m;
  }
  uint32_t w[8];
  load_chunk_words(bytes, base, n, w);
  ByteAtGlobal at{bytes, n};
  return classify_chunk(w, base, n, at, cls_tbl, KIND);
}

struct DsAtGlobal {
  const uint32_t* __restrict__ doc_bits;
  int64_t n_chunks;
  __device__ __forceinline__ uint32_t operator()(int64_t k) const { return (k >= 0 && k < n_chunks) ? __ldg(doc_bits + k) : 0u; }
};

// page summary packing: [11:0] chars, [23:12] kept splits, [35:24] chars since last doc start, [47:36] splits since, [48] has doc start
__device__ __forceinline__ uint64_t pack_sum(uint32_t tot, uint32_t aft, uint32_t flag) {
  // tot / aft carry two 16-bit fields each (chars | splits << 16)
  return (uint64_t)(tot & 0xFFFu) | ((uint64_t)((tot >> 16) & 0xFFFu) << 12) | ((uint64_t)(aft & 0xFFFu) << 24) |
         ((uint64_t)((aft >> 16) & 0xFFFu) << 36) | ((uint64_t)(flag & 1u) << 48);
}

// K1.  Each block owns a contiguous range of tiles (TC chunks = TC * 32 bytes each) and software-pipelines them:
// phase A of tile t+1 (class masks -> shared memory) runs before phase B of tile t (boundaries from the 64-byte
// windows), so the halo chunk on the right comes for free and the one on the left is kept from the previous tile.
//   A1  per lane: 2 x 16-byte loads, ASCII masks by SWAR (ascii_masks), store to shared memory
//   A2  per warp: the non-ASCII characters of the warp's 1 KB are compacted into a list and classified by all 32
//       lanes together (decode + 2-bit class table), results OR-ed into the owners' masks with shared atomics --
//       no lane diverges on "my chunk has 10 Cyrillic letters and yours has none"
//   B   per lane: windows from the neighbours' masks, boundary algebra (pretok_logic.cuh), 4-byte store of the bitmap
//       word, page summaries by warp reductions
#ifndef B2T_K1_MINBLOCKS
#define B2T_K1_MINBLOCKS 4
#endif
template <int KIND, int TC>
__global__ void __launch_bounds__(TC, B2T_K1_MINBLOCKS) pretok_scan_kernel(const uint8_t* __restrict__ bytes, int64_t n,
                                                         const uint32_t* __restrict__ doc_bits,
                                                         const uint32_t* __restrict__ cls_tbl,
                                                         uint32_t* __restrict__ start_bits, uint32_t* __restrict__ drop_bits,
                                                         uint64_t* __restrict__ page_sum, int64_t n_tiles, int64_t tiles_per_block, uint32_t one) {
  constexpr int NWARPS = TC / 32;
  __shared__ ChunkMasks sm[2][TC];
  __shared__ ChunkMasks sm_prev;                 // last chunk of the tile before the one in phase B
  __shared__ uint32_t s_spill[2][NWARPS][4];     // class bits spilling from a warp's last chunk into the next warp's first
  __shared__ uint16_t s_list[NWARPS][512];       // positions (within the warp's 1 KB) of its non-ASCII lead bytes
  __shared__ uint32_t s_wsum[NWARPS * 3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_chunks = n / CHUNK + 1;
  ByteAtGlobal at{bytes, n};
  DsAtGlobal dsat{doc_bits, n_chunks};
  const int64_t t_lo = (int64_t)blockIdx.x * tiles_per_block;
  const int64_t t_hi = (t_lo + tiles_per_block < n_tiles) ? t_lo + tiles_per_block : n_tiles;
  if (t_lo >= t_hi) return;

  // ---- phase A of one tile into buffer `buf`
  auto phase_a = [&](int64_t tile, int buf) {
    const int64_t c = tile * TC + tid, base = c * CHUNK;
    ChunkMasks m;
    uint32_t hi = 0, cont = 0;
    if (base < n) {
      uint32_t w[8];
      load_chunk_words(bytes, base, n, w);
      ascii_masks(KIND, w, m, &hi, &cont, one);
      if (base + CHUNK > n) {
        const uint32_t valid = 0xFFFFFFFFu >> (32 - (int)(n - base));
        m.lead &= valid; hi &= valid; cont &= valid;
      }
    } else { m.lead = m.L = m.N = m.S = m.SP = m.NL = m.AP = 0u; }
    sm[buf][tid] = m;
    // my OUTGOING spill slot (the incoming one may be written by the warp below at any time; slot 0 stays zero)
    if (lane < 4 && warp + 1 < NWARPS) s_spill[buf][warp + 1][lane] = 0u;
    __syncwarp();
    if (__any_sync(0xFFFFFFFFu, hi != 0u)) {
      // A2: compact the warp's non-ASCII lead bytes, then classify them with all lanes
      uint32_t nl = hi & ~cont;
      const int cnt = __popc(nl);
      int inc = cnt;
#pragma unroll
      for (int st = 1; st < 32; st <<= 1) { int o = __shfl_up_sync(0xFFFFFFFFu, inc, st); if (lane >= st) inc += o; }
      const int total = __shfl_sync(0xFFFFFFFFu, inc, 31);
      int off = inc - cnt;
      while (nl) { s_list[warp][off++] = (uint16_t)(lane * 32 + __ffs((int)nl) - 1); nl &= nl - 1u; }
      // a character that started in the previous TILE owns my first bytes (tid 0 only; inside the tile the owner
      // of the lead byte spills its bits into the neighbour below)
      if (tid == 0 && (cont & 1u)) {
        int back = 1;
        while (back < 3 && (at(base - back) & 0xC0u) == 0x80u) ++back;
        int len;
        const int64_t q = base - back;
        const uint32_t cl = decode_class(at(q), at(q + 1), at(q + 2), at(q + 3), cls_tbl, &len);
        const int cover = len - back;
        if (cover > 0 && cl != CLS_O) atomicOr(reinterpret_cast<uint32_t*>(&sm[buf][0]) + cl, (1u << cover) - 1u);
      }
      __syncwarp();
      const uint32_t wbase = (uint32_t)((tile * TC + warp * 32) * CHUNK);   // batches are < 2^31 bytes
      const uint32_t* __restrict__ words = reinterpret_cast<const uint32_t*>(bytes);
      uint32_t* smw = reinterpret_cast<uint32_t*>(&sm[buf][warp * 32]);
      constexpr int STRIDE = sizeof(ChunkMasks) / 4;
      for (int i = lane; i < total; i += 32) {
        const uint32_t idx = s_list[warp][i];
        const uint32_t gp = wbase + idx;
        const uint32_t a0 = __ldg(words + (gp >> 2)), a1 = ((gp | 3u) + 1u < (uint32_t)n) ? __ldg(words + (gp >> 2) + 1) : 0u;
        const uint32_t v = __funnelshift_r(a0, a1, (gp & 3u) * 8u);
        const uint32_t b0 = v & 0xFFu;
        const int len = 2 + (b0 >= 0xE0u) + (b0 >= 0xF0u);
        // branch-free decode (the three lengths are mixed inside a warp): the low 6 bits of all four bytes as if the
        // character had 4 bytes, shifted down by the bytes it does not have, lead-byte marker bits masked off
        const uint32_t t24 = ((v & 0x3Fu) << 18) | ((v << 4) & 0x3F000u) | ((v >> 10) & 0xFC0u) | ((v >> 24) & 0x3Fu);
        uint32_t cp = (t24 >> (6 * (4 - len))) & ((2u << (5 * len)) - 1u);
        cp = cp < 0x110000u ? cp : 0x10FFFFu;
        const uint32_t cl = (__ldg(cls_tbl + (cp >> 4)) >> ((cp & 15u) * 2u)) & 3u;
        if (cl != CLS_O) {
          const uint32_t owner = idx >> 5, p = idx & 31u, m = (1u << len) - 1u;
          atomicOr(smw + owner * STRIDE + cl, m << p);
          if (p + (uint32_t)len > 32u) {
            const uint32_t hi32 = m >> (32u - p);
            if (owner < 31u) atomicOr(smw + (owner + 1u) * STRIDE + cl, hi32);
            else if (warp + 1 < NWARPS) atomicOr(&s_spill[buf][warp + 1][cl], hi32);
          }
        }
      }
    }
  };
  auto load_masks = [&](int buf, int idx) -> ChunkMasks {
    ChunkMasks m = sm[buf][idx];
    if ((idx & 31) == 0) { const uint32_t* sp = s_spill[buf][idx >> 5]; m.L |= sp[CLS_L]; m.N |= sp[CLS_N]; m.S |= sp[CLS_S]; }
    return m;
  };

  // ---- prologue
  if (tid < 8) s_spill[tid >> 2][0][tid & 3] = 0u;
  if (tid == 0) {
    sm_prev = classify_global<KIND>(bytes, n, t_lo * TC - 1, cls_tbl);
  }
  phase_a(t_lo, (int)(t_lo & 1));

  for (int64_t tile = t_lo; tile < t_hi; ++tile) {
    const int buf = (int)(tile & 1), nbuf = buf ^ 1;
    if (tile + 1 < t_hi) phase_a(tile + 1, nbuf);
    else if (tid == 0) {  // first chunk after my range (phase B of my last tile looks 16 bytes into it)
      sm[nbuf][0] = classify_global<KIND>(bytes, n, (tile + 1) * TC, cls_tbl);
    }
    __syncthreads();
    // ---- phase B
    const int64_t c = tile * TC + tid;
    BoundaryOut r;
    ChunkMasks o = load_masks(buf, tid);
    // doc-start words of chunks c-1, c, c+1 (chunk indices fit 32 bits: batches are < 2^31 bytes; c - 1 wraps to "outside")
    auto ds32 = [&](uint32_t k) -> uint32_t { return k < (uint32_t)n_chunks ? __ldg(doc_bits + k) : 0u; };
    const uint32_t own_ds = ds32((uint32_t)c);
    {
      const ChunkMasks p = tid == 0 ? sm_prev : load_masks(buf, tid - 1);
      const ChunkMasks x = tid == TC - 1 ? load_masks(nbuf, 0) : load_masks(buf, tid + 1);
      Window w;
      w.lead = win(p.lead, o.lead, x.lead); w.L = win(p.L, o.L, x.L); w.N = win(p.N, o.N, x.N); w.S = win(p.S, o.S, x.S);
      w.SP = win(p.SP, o.SP, x.SP); w.NL = KIND == PT_LLAMA3 ? win(p.NL, o.NL, x.NL) : 0ull; w.AP = win(p.AP, o.AP, x.AP);
      w.DS = win(ds32((uint32_t)c - 1u), own_ds, ds32((uint32_t)c + 1u));
      const int64_t wb = c * CHUNK - 16;
      if (KIND == PT_GPT2) {
        r = boundaries_gpt2(w, wb, at);
      } else if (KIND == PT_LLAMA3) {
        LlamaCarry cy; cy.n_count_before_window = 0; cy.zone_before_window = false; cy.tail_after_window = false;
        r = boundaries_llama3(w, wb, at, cy);
        if (r.slow) {
          // rare: a run reaches beyond the window; the carry walks chunk masks re-classified from global memory
          struct GlobalMaskAt {
            const uint8_t* bytes; int64_t n; const uint32_t* cls;
            __device__ __forceinline__ ChunkMasks operator()(int64_t k) const { return classify_global<KIND>(bytes, n, k, cls); }
          } masks{bytes, n, cls_tbl};
          cy = llama_carry(c, n_chunks, masks, dsat);
          r = boundaries_llama3(w, wb, at, cy);
        }
      } else if (KIND == PT_WHITESPACE) {
        r = boundaries_whitespace(w);
      } else {
        r.start = (uint32_t)((w.DS & w.lead) >> 16); r.drop = 0; r.slow = 0;
      }
    }
    if (c < n_chunks) {
      start_bits[c] = r.start;
      if (KIND == PT_WHITESPACE) drop_bits[c] = r.drop;
    }
    // ---- page summaries (segmented: counts restart at the last doc start of the page), warp reductions
    {
      const uint32_t kept = r.start & ~r.drop;
      const uint32_t tot = (uint32_t)__popc(o.lead) | ((uint32_t)__popc(kept) << 16);
      const unsigned dsm = __ballot_sync(0xFFFFFFFFu, own_ds != 0u);
      const uint32_t wtot = __reduce_add_sync(0xFFFFFFFFu, tot);
      uint32_t waft = 0;
      if (dsm) {
        const int last = 31 - __clz((int)dsm);
        uint32_t mine = 0;
        if (lane > last) mine = tot;
        else if (lane == last) {
          const uint32_t from = ~bits_below(31 - __clz((int)own_ds));
          mine = (uint32_t)__popc(o.lead & from) | ((uint32_t)__popc(kept & from) << 16);
        }
        waft = __reduce_add_sync(0xFFFFFFFFu, mine);
      }
      if (lane == 0) { s_wsum[warp * 3] = wtot; s_wsum[warp * 3 + 1] = waft; s_wsum[warp * 3 + 2] = dsm != 0u; }
    }
    __syncthreads();  // all reads of sm[buf] / sm_prev are done; s_wsum is complete
    {
      constexpr int WPP = PAGE_CHUNKS / 32;  // warps per page (2)
      if (tid < NWARPS / WPP) {
        uint32_t t0 = s_wsum[(tid * WPP) * 3], a0 = s_wsum[(tid * WPP) * 3 + 1], f0 = s_wsum[(tid * WPP) * 3 + 2];
#pragma unroll
        for (int k = 1; k < WPP; ++k) {
          uint32_t t1 = s_wsum[(tid * WPP + k) * 3], a1 = s_wsum[(tid * WPP + k) * 3 + 1], f1 = s_wsum[(tid * WPP + k) * 3 + 2];
          a0 = f1 ? a1 : a0 + t1; t0 += t1; f0 |= f1;
        }
        const int64_t page = tile * (TC / PAGE_CHUNKS) + tid;
        if (page * PAGE <= n) page_sum[page] = pack_sum(t0, a0, f0);
      }
      if (tid == TC - 1) sm_prev = o;
    }
    // the next iteration's phase A writes sm[buf] and s_wsum is rewritten only after the next barrier
  }
}

// ------------------------------------------------------------------------------------------------ K1b
// Exclusive SEGMENTED scan over the page summaries (segments restart at document starts) -> page_carry.
// Two tiny kernels: (1) every block scans 1024 pages (one per thread, coalesced) and emits its block summary,
// (2) one block scans the block summaries.  A page whose block has no document start before it adds its block's
// carry when it is consumed (bit 31 of the low word says "already absolute").
__device__ __forceinline__ void sum_unpack(uint64_t s, uint32_t& tc, uint32_t& ts, uint32_t& ac, uint32_t& as, uint32_t& f) {
  tc = (uint32_t)(s & 0xFFFu); ts = (uint32_t)((s >> 12) & 0xFFFu); ac = (uint32_t)((s >> 24) & 0xFFFu);
  as = (uint32_t)((s >> 36) & 0xFFFu); f = (uint32_t)((s >> 48) & 1u);
}

struct SegVal { uint32_t f, c, s; };  // "carry after" as a function of "carry before": f ? (c, s) : before + (c, s)
__device__ __forceinline__ SegVal seg_combine(SegVal earlier, SegVal later) {
  SegVal r;
  r.f = earlier.f | later.f;
  r.c = later.f ? later.c : earlier.c + later.c;
  r.s = later.f ? later.s : earlier.s + later.s;
  return r;
}
// inclusive block scan (1024 threads); returns this thread's inclusive value, *excl = exclusive value
__device__ __forceinline__ SegVal seg_block_scan(SegVal v, SegVal* excl, SegVal* s_w /*32*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int st = 1; st < 32; st <<= 1) {
    SegVal o;
    o.f = __shfl_up_sync(0xFFFFFFFFu, v.f, st); o.c = __shfl_up_sync(0xFFFFFFFFu, v.c, st); o.s = __shfl_up_sync(0xFFFFFFFFu, v.s, st);
    if (lane >= st) v = seg_combine(o, v);
  }
  if (lane == 31) s_w[warp] = v;
  __syncthreads();
  if (warp == 0) {
    SegVal w = s_w[lane];
#pragma unroll
    for (int st = 1; st < 32; st <<= 1) {
      SegVal o;
      o.f = __shfl_up_sync(0xFFFFFFFFu, w.f, st); o.c = __shfl_up_sync(0xFFFFFFFFu, w.c, st); o.s = __shfl_up_sync(0xFFFFFFFFu, w.s, st);
      if (lane >= st) w = seg_combine(o, w);
    }
    s_w[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  SegVal before; before.f = 0; before.c = 0; before.s = 0;
  if (warp > 0) before = s_w[warp - 1];
  SegVal incl = seg_combine(before, v);
  SegVal pv;
  pv.f = __shfl_up_sync(0xFFFFFFFFu, v.f, 1); pv.c = __shfl_up_sync(0xFFFFFFFFu, v.c, 1); pv.s = __shfl_up_sync(0xFFFFFFFFu, v.s, 1);
  *excl = lane == 0 ? before : seg_combine(before, pv);
  __syncthreads();
  return incl;
}

constexpr int SCAN_BLOCK = 1024;

__global__ void __launch_bounds__(SCAN_BLOCK) page_scan_block_kernel(const uint64_t* __restrict__ page_sum, uint64_t* __restrict__ page_carry,
                                                                    uint64_t* __restrict__ block_sum, int64_t n_pages) {
  __shared__ SegVal s_w[32];
  const int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  SegVal v; v.f = 0; v.c = 0; v.s = 0;
  if (i < n_pages) {
    uint32_t tc, ts, ac, as, f;
    sum_unpack(__ldg(page_sum + i), tc, ts, ac, as, f);
    v.f = f; v.c = f ? ac : tc; v.s = f ? as : ts;
  }
  SegVal ex;
  SegVal incl = seg_block_scan(v, &ex, s_w);
  if (i < n_pages) page_carry[i] = (uint64_t)(ex.c | (ex.f << 31)) | ((uint64_t)ex.s << 32);
  if (threadIdx.x == SCAN_BLOCK - 1) block_sum[blockIdx.x] = (uint64_t)(incl.c | (incl.f << 31)) | ((uint64_t)incl.s << 32);
}

__global__ void __launch_bounds__(SCAN_BLOCK) page_scan_top_kernel(const uint64_t* __restrict__ block_sum, uint64_t* __restrict__ block_carry,
                                                                  int64_t n_blocks) {
  __shared__ SegVal s_w[32];
  __shared__ SegVal s_run;
  if (threadIdx.x == 0) { s_run.f = 0; s_run.c = 0; s_run.s = 0; }
  __syncthreads();
  for (int64_t base = 0; base < n_blocks; base += SCAN_BLOCK) {
    const int64_t i = base + threadIdx.x;
    SegVal v; v.f = 0; v.c = 0; v.s = 0;
    if (i < n_blocks) { uint64_t x = __ldg(block_sum + i); v.f = (uint32_t)(x >> 31) & 1u; v.c = (uint32_t)x & 0x7FFFFFFFu; v.s = (uint32_t)(x >> 32); }
    SegVal ex;
    SegVal incl = seg_block_scan(v, &ex, s_w);
    const SegVal run = s_run;
    const SegVal ex_abs = seg_combine(run, ex);
    if (i < n_blocks) block_carry[i] = (uint64_t)ex_abs.c | ((uint64_t)ex_abs.s << 32);
    __syncthreads();
    if (threadIdx.x == SCAN_BLOCK - 1) s_run = seg_combine(run, incl);
    __syncthreads();
  }
}

}  // namespace b2t

// ================================================================================================ K1, streaming form
// pretok_stream_kernel: the production K1.  Every WARP streams through its own contiguous range of the batch, 1 KB
// (32 chunks of 32 bytes, one per lane) per iteration, and is independent of all other warps: no shared memory, no
// block barrier.  Iteration i+1 is classified (pretok_fast.cuh: bit planes -> classes) before the boundaries of
// iteration i are evaluated, so a lane gets everything it needs from its neighbours with warp shuffles:
//   the chunk before lane 0   = lane 31 of the previous iteration (still in that lane's registers),
//   the chunk after lane 31   = lane 0 of the next iteration (already classified).
// The range ends are handled by classifying one extra KB on either side.  Page summaries (2 KB = two iterations of the
// same warp) are combined in registers.  Algorithmic traffic is unchanged: 1.25 B per input byte.
#include "pretok_fast.cuh"

namespace b2t {

// Exact window code (pretok_logic.cuh) for one chunk, everything re-read from global memory.  Used by the rare
// fallback of the fast GPT-2 algebra and as the Llama-3 slow path.
template <int KIND>
__device__ __noinline__ uint32_t exact_chunk_start(const uint8_t* __restrict__ bytes, int64_t n, int64_t c,
                                                  const uint32_t* __restrict__ cls_tbl, const uint32_t* __restrict__ doc_bits) {
  const int64_t n_chunks = n / CHUNK + 1;
  ByteAtGlobal at{bytes, n};
  DsAtGlobal dsat{doc_bits, n_chunks};
  const ChunkMasks p = classify_global<KIND>(bytes, n, c - 1, cls_tbl), o = classify_global<KIND>(bytes, n, c, cls_tbl),
                   x = classify_global<KIND>(bytes, n, c + 1, cls_tbl);
  Window w;
  w.lead = win(p.lead, o.lead, x.lead); w.L = win(p.L, o.L, x.L); w.N = win(p.N, o.N, x.N); w.S = win(p.S, o.S, x.S);
  w.SP = win(p.SP, o.SP, x.SP); w.NL = KIND == PT_LLAMA3 ? win(p.NL, o.NL, x.NL) : 0ull; w.AP = win(p.AP, o.AP, x.AP);
  w.DS = win(dsat(c - 1), dsat(c), dsat(c + 1));
  const int64_t wb = c * CHUNK - 16;
  if (KIND == PT_GPT2) return boundaries_gpt2(w, wb, at).start;
  if (KIND == PT_LLAMA3) {
    struct GlobalMaskAt {
      const uint8_t* bytes; int64_t n; const uint32_t* cls;
      __device__ __forceinline__ ChunkMasks operator()(int64_t k) const { return classify_global<KIND>(bytes, n, k, cls); }
    } masks{bytes, n, cls_tbl};
    const LlamaCarry cy = llama_carry(c, n_chunks, masks, dsat);
    return boundaries_llama3(w, wb, at, cy).start;
  }
  return boundaries_whitespace(w).start;
}

#ifndef B2T_K1S_THREADS
#define B2T_K1S_THREADS 128
#endif
#ifndef B2T_K1S_MINBLOCKS
#define B2T_K1S_MINBLOCKS 8
#endif
#ifndef B2T_K1W_MINBLOCKS
#define B2T_K1W_MINBLOCKS 4   // window (Llama-3) form: more live state
#endif

#ifdef B2T_K1_DEBUG
__device__ uint32_t g_k1_dbg[8192 * 8];
#endif
// byte access with 32-bit positions (batches are < 2^31 bytes; "negative" positions wrap to out-of-range)
struct ByteAt32 {
  const uint8_t* __restrict__ p;
  uint32_t n;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return i < n ? (uint32_t)__ldg(p + i) : 0u; }
};
// the four bytes at q..q+3 (q < n) as one word: two aligned loads and a funnel shift
struct At4Global {
  const uint32_t* __restrict__ words;
  uint32_t n;
  __device__ __forceinline__ uint32_t operator()(uint32_t q) const {
    const uint32_t a0 = __ldg(words + (q >> 2)), a1 = ((q | 3u) + 1u < n) ? __ldg(words + (q >> 2) + 1) : 0u;
    return __funnelshift_r(a0, a1, (q & 3u) * 8u);
  }
};

// the chunk that holds the end of the batch: aligned word loads, bytes at or past n masked off (stays in registers)
__device__ __forceinline__ void load_tail_words(const uint8_t* __restrict__ bytes, uint32_t base, uint32_t n, uint32_t w[8]) {
  const uint32_t* __restrict__ words = reinterpret_cast<const uint32_t*>(bytes + base);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t p = base + 4u * j;
    uint32_t x = 0u;
    if (p < n) {
      x = __ldg(words + j);
      if (n - p < 4u) x &= (1u << (8u * (n - p))) - 1u;
    }
    w[j] = x;
  }
}

// ---------------------------------------------------------------------------------------------- lean form (GPT-2, Whitespace, no regex)
// Stage A (iteration j): bit planes, classes, and all of the boundary algebra that looks backwards.  Stage B
// (iteration j - 1): the one bit that looks forwards (finalize_gpt2), the store and the page summary.  Between the two
// a lane keeps six words, so the kernel runs at 64 registers.  There is one copy of the code: the loop starts one
// iteration before the warp's range (its output is discarded, only the carries are kept).
// ADDED: the batch went through the added-token extraction (added_kernels.cuh): the regex sees hard_bits (document starts +
// span boundaries) where it otherwise sees doc_bits, splits inside a span are cleared (inner_bits), Whitespace never drops a
// span's bytes (added_bits | inner_bits).  The page summaries keep counting from the DOCUMENT starts.
struct AddedBits { const uint32_t* hard; const uint32_t* inner; const uint32_t* added; };
template <int KIND, bool ADDED = false>
__global__ void __launch_bounds__(B2T_K1S_THREADS, B2T_K1S_MINBLOCKS)
pretok_lean_kernel(const uint8_t* __restrict__ bytes, int64_t n64, const uint32_t* __restrict__ doc_bits,
                   const uint32_t* __restrict__ cls_tbl, uint32_t* __restrict__ start_bits, uint32_t* __restrict__ drop_bits,
                   uint64_t* __restrict__ page_sum, int n_kb, int kb_per_warp, SwapMasks masks, AddedBits ab = AddedBits{nullptr, nullptr, nullptr}) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int it_lo = gw * kb_per_warp;             // even: a page is two consecutive iterations of one warp
  if (it_lo >= n_kb) return;
  const int it_hi = it_lo + kb_per_warp < n_kb ? it_lo + kb_per_warp : n_kb;
  const uint32_t n = (uint32_t)n64;
  const uint32_t n_chunks = n / CHUNK + 1;
  const At4Global at4{reinterpret_cast<const uint32_t*>(bytes), n};
  const int up = (lane + 31) & 31, down = (lane + 1) & 31;

  auto load = [&](int it, uint32_t w[8]) {       // it = -1: nothing there
    const uint32_t base = ((uint32_t)it * 32u + lane) * CHUNK;
    if (it >= 0 && base + CHUNK <= n) {
      const uint4* q = reinterpret_cast<const uint4*>(bytes + base);
      const uint4 a = __ldg(q), b = __ldg(q + 1);
      w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = 0u;
      if (it >= 0 && base < n) load_tail_words(bytes, base, n, w);
    }
  };

  // carries of my chunk of the previous iteration (lane 31's chunk precedes lane 0's)
  uint32_t kL = 0u, kN = 0u, kS = 0u, kSP = 0u, k_ov = 0u;
  // stage A results of iteration j - 1, waiting for stage B
  uint32_t p_start = 0u, p_drop = 0u, p_lead = 0u, p_head = 0u, p_ds = 0u, p_dsn = 0u, p_s31 = 0u;
  uint32_t p_inner = 0u, p_doc = 0u;              // ADDED only
  uint32_t h_tot = 0u, h_aft = 0u, h_flag = 0u;   // first half of the current page

  uint32_t w[8];
  load(it_lo > 0 ? it_lo - 1 : -1, w);
#pragma unroll 1
  for (int j = it_lo - 1; j <= it_hi; ++j) {
    // ================= stage A: iteration j
    const uint32_t c = (uint32_t)j * 32u + lane, base = c * CHUNK;
    // interior iteration (warp-uniform): this KB and the first chunk of the next one lie inside the batch
    const bool interior = j >= 0 && ((uint32_t)j * 32u + 33u) * CHUNK <= n;
    uint32_t valid = 0xFFFFFFFFu;
    if (!interior) valid = j < 0 || base >= n ? 0u : (n - base >= CHUNK ? 0xFFFFFFFFu : (0xFFFFFFFFu >> (32u - (n - base))));
    uint32_t b[8];
    bitslice32(w, b, masks);
    if (j + 1 < it_hi && ((uint32_t)j * 32u + 64u) * CHUNK <= n) {   // prefetch; the common case first: a whole KB inside the batch
      const uint4* q = reinterpret_cast<const uint4*>(bytes + (uint32_t)(base + 32u * CHUNK));   // (j = -1: the sum wraps to lane * 32)
      const uint4 a = __ldg(q), bq = __ldg(q + 1);
      w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = bq.x; w[5] = bq.y; w[6] = bq.z; w[7] = bq.w;
    } else {
      load(j + 1 <= it_hi ? j + 1 : -1, w);      // range / batch ends (the KB after the range is classified too, for its first bytes)
    }
    FastCls m = classify_planes<KIND>(b, valid);
    if (__any_sync(FULL, m.hi != 0u)) {
      if (m.unc) resolve_uncertain(m, at4, base, cls_tbl);
      fill_own(m);
    }
    PrevTop pt;
    pt.L = __shfl_sync(FULL, lane == 31 ? kL : m.L, up);
    pt.S = __shfl_sync(FULL, lane == 31 ? kS : m.S, up);
    if (KIND == PT_GPT2) {
      pt.N = __shfl_sync(FULL, lane == 31 ? kN : m.N, up);
      pt.SP = __shfl_sync(FULL, lane == 31 ? kSP : m.SP, up);
    } else { pt.N = 0u; pt.SP = 0u; }
    if (m.cont & 1u) spill_in(m, pt.L, pt.N, pt.S);
    kL = m.L; kN = m.N; kS = m.S; kSP = m.SP;
    // ds / ds_next: where a new string starts as far as the regex is concerned (ADDED: hard_bits)
    const uint32_t* __restrict__ rbits = ADDED ? ab.hard : doc_bits;
    uint32_t ds, ds_next;
    if (interior) { ds = __ldg(rbits + c); ds_next = __ldg(rbits + c + 1); }
    else { ds = (j >= 0 && c < n_chunks) ? __ldg(rbits + c) : 0u; ds_next = (j >= 0 && c + 1 < n_chunks) ? __ldg(rbits + c + 1) : 0u; }
    uint32_t inner = 0u, keep = 0u, docw = ds;
    if (ADDED && j >= 0 && c < n_chunks) {
      inner = __ldg(ab.inner + c); docw = __ldg(doc_bits + c);
      if (pretok_drops_whitespace(KIND)) keep = inner | __ldg(ab.added + c);
    }
    uint32_t start, drop = 0u;
    if (KIND == PT_GPT2) {
      const FastOut o = fast_gpt2(m, pt, 1u, 1u, ds, ds_next, base, at4);
      Overflow in; in.bits = __shfl_sync(FULL, lane == 31 ? k_ov : o.ov.bits, up);
      k_ov = o.ov.bits;
      start = apply_overflow(o.start, m.lead, in);
    } else if (KIND == PT_WHITESPACE || KIND == PT_BERT) {
      const FastOut o = KIND == PT_BERT ? fast_bert(m, pt, ds) : fast_whitespace(m, pt, ds);
      start = o.start; drop = o.drop & ~keep;
    } else {
      start = ds & m.lead;
    }
    const uint32_t head = (m.lead & 15u) | ((m.S & 15u) << 4);
    // ================= stage B: iteration j - 1
    if (j > it_lo) {
      const int it = j - 1;
      const uint32_t pc = (uint32_t)it * 32u + lane;
      uint32_t fin = p_start;
      if (KIND == PT_GPT2) {
        const uint32_t xh = __shfl_sync(FULL, lane == 0 ? head : p_head, down);
        fin = finalize_gpt2(fin, p_lead, p_s31, xh & 15u, xh >> 4, p_dsn);
      }
      if (ADDED) fin &= ~p_inner;      // an added token's span is one pre-token
      const uint32_t p_docw = ADDED ? p_doc : p_ds;
      if (pc < n_chunks) {
        start_bits[pc] = fin;
        if (pretok_drops_whitespace(KIND)) drop_bits[pc] = p_drop;
      }
      // ---- page summary (segmented: counts restart at the last doc start of the page); one iteration is half a page
      const uint32_t kept = fin & ~p_drop;
      const uint32_t tot = (uint32_t)__popc(p_lead) | ((uint32_t)__popc(kept) << 16);
      const unsigned dsm = __ballot_sync(FULL, p_docw != 0u);
      const uint32_t wtot = __reduce_add_sync(FULL, tot);
      uint32_t waft = 0u;
      if (dsm) {
        const int last = 31 - __clz((int)dsm);
        uint32_t mine = 0u;
        if (lane > last) mine = tot;
        else if (lane == last) {
          const uint32_t from = ~bits_below(31 - __clz((int)p_docw));
          mine = (uint32_t)__popc(p_lead & from) | ((uint32_t)__popc(kept & from) << 16);
        }
        waft = __reduce_add_sync(FULL, mine);
      }
      const uint32_t wflag = dsm != 0u;
      const uint32_t page = (uint32_t)it >> 1;
      if (it & 1) {
        const uint32_t a = wflag ? waft : h_aft + wtot, t = h_tot + wtot, f = h_flag | wflag;
        if (lane == 0 && page * (uint32_t)PAGE <= n) page_sum[page] = pack_sum(t, a, f);
      } else {
        h_tot = wtot; h_aft = waft; h_flag = wflag;
        if (it + 1 >= n_kb && lane == 0 && page * (uint32_t)PAGE <= n) page_sum[page] = pack_sum(h_tot, h_aft, h_flag);  // the batch ends in the first half
      }
    }
    p_start = start; p_drop = drop; p_lead = m.lead; p_head = head; p_ds = ds; p_dsn = ds_next; p_s31 = m.S >> 31;
    if (ADDED) { p_inner = inner; p_doc = docw; }
  }
}

// ---------------------------------------------------------------------------------------------- window form (Llama-3)
// Same streaming structure with the 64-bit window algebra of pretok_logic.cuh (two iterations per trip).
template <int KIND, bool ADDED = false>
__global__ void __launch_bounds__(B2T_K1S_THREADS, B2T_K1W_MINBLOCKS)
pretok_stream_kernel(const uint8_t* __restrict__ bytes, int64_t n64, const uint32_t* __restrict__ doc_bits,
                     const uint32_t* __restrict__ cls_tbl, uint32_t* __restrict__ start_bits, uint32_t* __restrict__ drop_bits,
                     uint64_t* __restrict__ page_sum, int n_kb, int kb_per_warp, AddedBits ab = AddedBits{nullptr, nullptr, nullptr}) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int it_lo = gw * kb_per_warp;             // even: a page is two consecutive iterations of one warp
  if (it_lo >= n_kb) return;
  const int it_hi = it_lo + kb_per_warp < n_kb ? it_lo + kb_per_warp : n_kb;
  const uint32_t n = (uint32_t)n64;
  const uint32_t n_chunks = n / CHUNK + 1;
  static_assert(KIND == PT_LLAMA3, "the window form is the Llama-3 kernel");
  const At4Global at4{reinterpret_cast<const uint32_t*>(bytes), n};
  const int up = (lane + 31) & 31, down = (lane + 1) & 31;

  auto load = [&](int it, uint32_t w[8]) {       // it = -1: nothing there
    const uint32_t base = ((uint32_t)it * 32u + lane) * CHUNK;
    if (it >= 0 && base + CHUNK <= n) {
      const uint4* q = reinterpret_cast<const uint4*>(bytes + base);
      const uint4 a = __ldg(q), b = __ldg(q + 1);
      w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = 0u;
      if (it >= 0 && base < n) load_tail_words(bytes, base, n, w);   // the batch ends inside this chunk
    }
  };
  // classes of iteration `it`; `prev` = this lane's classes of the iteration before (lane 31's chunk precedes lane 0's)
  auto classify = [&](const uint32_t w[8], int it, const FastCls& prev, FastCls& m, FastCls& pm) {
    const uint32_t base = ((uint32_t)it * 32u + lane) * CHUNK;
    const uint32_t valid = it < 0 || base >= n ? 0u : (n - base >= CHUNK ? 0xFFFFFFFFu : (0xFFFFFFFFu >> (32u - (n - base))));
    uint32_t b[8];
    bitslice32(w, b);
    m = classify_planes<KIND>(b, valid);
#ifdef B2T_K1_NOANY
    {
#else
    if (__any_sync(FULL, m.hi != 0u)) {
#endif
      if (m.unc) resolve_uncertain(m, at4, base, cls_tbl);
      fill_own(m);
    }
    // the chunk before mine (only bit 31 of the class words matters to the fast algebra; Llama-3 uses the top halves)
    pm.L = __shfl_sync(FULL, lane == 31 ? prev.L : m.L, up);
    pm.N = KIND == PT_WHITESPACE ? 0u : __shfl_sync(FULL, lane == 31 ? prev.N : m.N, up);
    pm.S = __shfl_sync(FULL, lane == 31 ? prev.S : m.S, up);
    pm.SP = KIND == PT_WHITESPACE ? 0u : __shfl_sync(FULL, lane == 31 ? prev.SP : m.SP, up);
    if (KIND == PT_LLAMA3) {
      pm.lead = __shfl_sync(FULL, lane == 31 ? prev.lead : m.lead, up);
      pm.NL = __shfl_sync(FULL, lane == 31 ? prev.NL : m.NL, up);
      pm.AP = __shfl_sync(FULL, lane == 31 ? prev.AP : m.AP, up);
    }
    if (m.cont & 1u) spill_in(m, pm.L, pm.N, pm.S);
  };

  uint32_t h_tot = 0u, h_aft = 0u, h_flag = 0u;   // first half of the current page

  // One iteration: classify it+1 (words in w) into nxt / pnxt, refill w with the words of it+3, then boundaries,
  // stores and page summary of iteration `it` (classes in cur / pcur).
  auto step = [&](int it, const FastCls& cur, const FastCls& pcur, FastCls& nxt, FastCls& pnxt, uint32_t w[8]) {
    classify(w, it + 1, cur, nxt, pnxt);
    load(it + 3 <= it_hi ? it + 3 : -1, w);      // prefetch (the KB after the range is classified too, for its first bytes)

    const uint32_t c = (uint32_t)it * 32u + lane, base = c * CHUNK;
    const uint32_t* __restrict__ rbits = ADDED ? ab.hard : doc_bits;   // string starts as the regex sees them
    const uint32_t ds = c < n_chunks ? __ldg(rbits + c) : 0u;
    const uint32_t ds_next = c + 1 < n_chunks ? __ldg(rbits + c + 1) : 0u;
    const uint32_t docw = ADDED ? (c < n_chunks ? __ldg(doc_bits + c) : 0u) : ds;
    uint32_t start, drop = 0u;
    {
      // window algebra of pretok_logic.cuh on [16 B of the previous chunk | mine | 16 B of the next chunk]
      Window wd;
      auto nextw = [&](uint32_t mine, uint32_t theirs) { return __shfl_sync(FULL, lane == 0 ? theirs : mine, down); };
      wd.lead = win(pcur.lead, cur.lead, nextw(cur.lead, nxt.lead)); wd.L = win(pcur.L, cur.L, nextw(cur.L, nxt.L));
      wd.N = win(pcur.N, cur.N, nextw(cur.N, nxt.N)); wd.S = win(pcur.S, cur.S, nextw(cur.S, nxt.S));
      wd.SP = win(pcur.SP, cur.SP, nextw(cur.SP, nxt.SP)); wd.NL = win(pcur.NL, cur.NL, nextw(cur.NL, nxt.NL));
      wd.AP = win(pcur.AP, cur.AP, nextw(cur.AP, nxt.AP));
      const uint32_t ds_prev = (c >= 1u && c - 1u < n_chunks) ? __ldg(rbits + c - 1) : 0u;
      wd.DS = win(ds_prev, ds, ds_next);
      LlamaCarry cy; cy.n_count_before_window = 0; cy.zone_before_window = false; cy.tail_after_window = false;
      const ByteAtGlobal at64{bytes, n64};
      const BoundaryOut r = boundaries_llama3(wd, (int64_t)base - 16, at64, cy);
      start = r.start;
      if (r.slow) start = exact_chunk_start<KIND>(bytes, n64, c, cls_tbl, rbits);
    }
    if (ADDED && c < n_chunks) start &= ~__ldg(ab.inner + c);   // an added token's span is one pre-token
    if (c < n_chunks) {
      start_bits[c] = start;
      if (KIND == PT_WHITESPACE) drop_bits[c] = drop;
    }
#ifdef B2T_K1_DEBUG
    if (c < 8192u) { uint32_t* d = g_k1_dbg + c * 8; d[0] = cur.lead; d[1] = cur.cont; d[2] = cur.L; d[3] = cur.N; d[4] = cur.S; d[5] = cur.SP; d[6] = pcur.L; d[7] = start; }
#endif
    // ---- page summary (segmented: counts restart at the last doc start of the page); this iteration is half a page
    const uint32_t kept = start & ~drop;
    const uint32_t tot = (uint32_t)__popc(cur.lead) | ((uint32_t)__popc(kept) << 16);
    const unsigned dsm = __ballot_sync(FULL, docw != 0u);
    const uint32_t wtot = __reduce_add_sync(FULL, tot);
    uint32_t waft = 0u;
    if (dsm) {
      const int last = 31 - __clz((int)dsm);
      uint32_t mine = 0u;
      if (lane > last) mine = tot;
      else if (lane == last) {
        const uint32_t from = ~bits_below(31 - __clz((int)docw));
        mine = (uint32_t)__popc(cur.lead & from) | ((uint32_t)__popc(kept & from) << 16);
      }
      waft = __reduce_add_sync(FULL, mine);
    }
    const uint32_t wflag = dsm != 0u;
    const uint32_t page = (uint32_t)it >> 1;
    if (it & 1) {
      const uint32_t a = wflag ? waft : h_aft + wtot, t = h_tot + wtot, f = h_flag | wflag;
      if (lane == 0 && page * (uint32_t)PAGE <= n) page_sum[page] = pack_sum(t, a, f);
    } else {
      h_tot = wtot; h_aft = waft; h_flag = wflag;
      if (it + 1 >= n_kb && lane == 0 && page * (uint32_t)PAGE <= n) page_sum[page] = pack_sum(h_tot, h_aft, h_flag);  // the batch ends in the first half
    }
  };

  FastCls A, pA, B, pB;   // p*: the chunk before (as far as needed)
  uint32_t w0[8], w1[8];
  {
    FastCls zero;
    zero.lead = zero.cont = zero.hi = zero.L = zero.N = zero.S = zero.SP = zero.AP = zero.NL = zero.unc = 0u;
    load(it_lo - 1, w0);
    classify(w0, it_lo - 1, zero, B, pB);
    load(it_lo, w0);
    classify(w0, it_lo, B, A, pA);
    load(it_lo + 1, w0);
    load(it_lo + 2 <= it_hi ? it_lo + 2 : -1, w1);
  }
  // two iterations per trip, the class registers ping-pong (A holds the even iteration, B the odd one)
  for (int it = it_lo; it < it_hi; it += 2) {
    step(it, A, pA, B, pB, w0);
    if (it + 1 < it_hi) step(it + 1, B, pB, A, pA, w1);
  }
}

}  // namespace b2t
