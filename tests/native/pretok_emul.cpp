// tests/native/pretok_emul.cpp -- runs the device pre-tokenization mask logic (tokenizers_b200/csrc/pretok_logic.cuh)
// on the CPU, chunk by chunk, exactly as the CUDA kernel composes it.  TEST INFRASTRUCTURE: lets the CPU test suite
// fuzz the bit-parallel boundary predicates against the oracle without a GPU.  Not linked into the product library.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../tokenizers_b200/csrc/pretok_logic.cuh"

using namespace b2t;

extern "C" int b2t_emul_pretok(int kind, const uint8_t* bytes, uint64_t n, const uint64_t* doc_off, uint32_t n_docs,
                               const uint32_t* cls_tbl, uint32_t* start_bits, uint32_t* drop_bits) {
  int64_t n_chunks = (int64_t)(n / CHUNK) + 1;
  std::vector<ChunkMasks> M(n_chunks);
  std::vector<uint32_t> DS(n_chunks + 1, 0);
  for (uint32_t d = 0; d <= n_docs; ++d) DS[doc_off[d] / 32] |= 1u << (doc_off[d] % 32);
  auto at = [&](int64_t p) -> uint32_t { return (p >= 0 && (uint64_t)p < n) ? bytes[p] : 0u; };
  for (int64_t c = 0; c < n_chunks; ++c) {
    uint32_t w[8];
    for (int j = 0; j < 8; ++j) {
      uint32_t x = 0;
      for (int b = 0; b < 4; ++b) x |= at(c * 32 + j * 4 + b) << (8 * b);
      w[j] = x;
    }
    M[c] = classify_chunk(w, c * 32, (int64_t)n, at, cls_tbl, kind);
  }
  ChunkMasks zero;
  std::memset(&zero, 0, sizeof(zero));
  auto masks = [&](int64_t k) -> ChunkMasks { return (k >= 0 && k < n_chunks) ? M[k] : zero; };
  auto dsat = [&](int64_t k) -> uint32_t { return (k >= 0 && k < n_chunks) ? DS[k] : 0u; };
  for (int64_t c = 0; c < n_chunks; ++c) {
    ChunkMasks p = masks(c - 1), o = masks(c), x = masks(c + 1);
    Window w;
    w.lead = win(p.lead, o.lead, x.lead); w.L = win(p.L, o.L, x.L); w.N = win(p.N, o.N, x.N); w.S = win(p.S, o.S, x.S);
    w.SP = win(p.SP, o.SP, x.SP); w.NL = win(p.NL, o.NL, x.NL); w.AP = win(p.AP, o.AP, x.AP);
    w.DS = win(dsat(c - 1), dsat(c), dsat(c + 1));
    int64_t wb = c * 32 - 16;
    BoundaryOut r;
    if (kind == PT_GPT2) r = boundaries_gpt2(w, wb, at);
    else if (kind == PT_LLAMA3) {
      LlamaCarry cy; cy.n_count_before_window = 0; cy.zone_before_window = false; cy.tail_after_window = false;
      r = boundaries_llama3(w, wb, at, cy);
      if (r.slow) { cy = llama_carry(c, n_chunks, masks, dsat); r = boundaries_llama3(w, wb, at, cy); }
    } else if (kind == PT_WHITESPACE) r = boundaries_whitespace(w);
    else { r.start = (uint32_t)((w.DS & w.lead) >> 16); r.drop = 0; r.slow = 0; }
    start_bits[c] = r.start;
    if (drop_bits) drop_bits[c] = r.drop;
  }
  return 0;
}
