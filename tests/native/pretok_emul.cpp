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

// ------------------------------------------------------------------------------------------------ fast path (pretok_fast.cuh)
#include "../../tokenizers_b200/csrc/pretok_fast.cuh"

// The streaming kernel's composition, chunk by chunk: classification (bit planes, certain classes, table look-ups, fill,
// spill-in from the previous chunk), the 32-bit boundary algebra with the next chunk's first-byte bits, contraction
// overflow into the next chunk, and the exact window code for chunks that ask for the fallback.
// fallbacks_out (optional): number of chunks that took the fallback.
template <int KIND>
static void emul_fast(const uint8_t* bytes, uint64_t n, const uint64_t* doc_off, uint32_t n_docs, const uint32_t* cls_tbl,
                      uint32_t* start_bits, uint32_t* drop_bits, uint64_t* fallbacks_out, uint32_t* planes_out = nullptr) {
  const int64_t n_chunks = (int64_t)(n / CHUNK) + 1;
  std::vector<uint32_t> DS(n_chunks + 2, 0);
  for (uint32_t d = 0; d <= n_docs; ++d) DS[doc_off[d] / 32] |= 1u << (doc_off[d] % 32);
  auto at = [&](int64_t p) -> uint32_t { return (p >= 0 && (uint64_t)p < n) ? bytes[p] : 0u; };
  auto at4 = [&](int64_t p) -> uint32_t { return at(p) | (at(p + 1) << 8) | (at(p + 2) << 16) | (at(p + 3) << 24); };
  std::vector<FastCls> M(n_chunks + 1);
  std::vector<PrevTop> PT(n_chunks + 1);
  FastCls prev;
  std::memset(&prev, 0, sizeof(prev));
  for (int64_t c = 0; c <= n_chunks; ++c) {
    uint32_t w[8], b[8];
    for (int j = 0; j < 8; ++j) {
      uint32_t x = 0;
      for (int k = 0; k < 4; ++k) x |= at(c * 32 + j * 4 + k) << (8 * k);
      w[j] = x;
    }
    const int64_t base = c * 32;
    const uint32_t valid = base + 32 <= (int64_t)n ? 0xFFFFFFFFu : (base >= (int64_t)n ? 0u : (0xFFFFFFFFu >> (32 - (int)((int64_t)n - base))));
    bitslice32(w, b);
    FastCls m = classify_planes<KIND>(b, valid);
    if (m.unc) resolve_uncertain(m, at4, base, cls_tbl);
    fill_own(m);
    PrevTop pt; pt.L = prev.L; pt.N = prev.N; pt.S = prev.S; pt.SP = prev.SP;
    if (m.cont & 1u) spill_in(m, pt.L, pt.N, pt.S);
    M[c] = m; PT[c] = pt; prev = m;
  }
  // exact window code (pretok_logic.cuh) for the fallback
  auto exact = [&](int64_t c, uint32_t* drop) -> uint32_t {
    auto masks = [&](int64_t k) -> ChunkMasks {
      ChunkMasks z; std::memset(&z, 0, sizeof(z));
      if (k < 0 || k >= n_chunks) return z;
      uint32_t w[8];
      for (int j = 0; j < 8; ++j) { uint32_t x = 0; for (int q = 0; q < 4; ++q) x |= at(k * 32 + j * 4 + q) << (8 * q); w[j] = x; }
      return classify_chunk(w, k * 32, (int64_t)n, at, cls_tbl, KIND);
    };
    auto dsat = [&](int64_t k) -> uint32_t { return (k >= 0 && k < n_chunks) ? DS[k] : 0u; };
    ChunkMasks p = masks(c - 1), o = masks(c), x = masks(c + 1);
    Window w;
    w.lead = win(p.lead, o.lead, x.lead); w.L = win(p.L, o.L, x.L); w.N = win(p.N, o.N, x.N); w.S = win(p.S, o.S, x.S);
    w.SP = win(p.SP, o.SP, x.SP); w.NL = win(p.NL, o.NL, x.NL); w.AP = win(p.AP, o.AP, x.AP);
    w.DS = win(dsat(c - 1), dsat(c), dsat(c + 1));
    BoundaryOut r;
    if (KIND == PT_GPT2) r = boundaries_gpt2(w, c * 32 - 16, at);
    else r = boundaries_whitespace(w);
    *drop = r.drop;
    return r.start;
  };
  Overflow ov_in; ov_in.bits = 0u;
  uint64_t nfb = 0;
  for (int64_t c = 0; c < n_chunks; ++c) {
    const FastCls& m = M[c];
    const FastCls& x = M[c + 1];
    FastOut o;
    if (KIND == PT_GPT2) o = fast_gpt2(m, PT[c], 1u, 1u, DS[c], DS[c + 1], c * 32, at4);
    else if (KIND == PT_WHITESPACE) o = fast_whitespace(m, PT[c], DS[c]);
    else if (KIND == PT_BERT) o = fast_bert(m, PT[c], DS[c]);
    else { o.start = DS[c] & m.lead; o.drop = 0; o.fallback = 0; o.ov.bits = 0; }
    uint32_t start = apply_overflow(o.start, m.lead, ov_in), drop = o.drop;
    if (KIND == PT_GPT2) {
      const uint32_t before = start;
      start = finalize_gpt2(start, m.lead, m.S >> 31, x.lead & 15u, x.S & 15u, DS[c + 1]);
      if (start != before && (m.hi >> 31)) ++nfb;   // counts the straddling multi-byte spaces that were given back
    }
    (void)exact;
    start_bits[c] = start;
    if (drop_bits) drop_bits[c] = drop;
    if (planes_out) { uint32_t* d = planes_out + c * 8; d[0] = m.lead; d[1] = m.cont; d[2] = m.L; d[3] = m.N; d[4] = m.S; d[5] = m.SP; d[6] = PT[c].L; d[7] = start; }
    ov_in = o.ov;
  }
  if (fallbacks_out) *fallbacks_out = nfb;
}

extern "C" int b2t_emul_pretok_fast(int kind, const uint8_t* bytes, uint64_t n, const uint64_t* doc_off, uint32_t n_docs,
                                    const uint32_t* cls_tbl, uint32_t* start_bits, uint32_t* drop_bits, uint64_t* fallbacks) {
  if (kind == PT_GPT2) emul_fast<PT_GPT2>(bytes, n, doc_off, n_docs, cls_tbl, start_bits, drop_bits, fallbacks);
  else if (kind == PT_WHITESPACE) emul_fast<PT_WHITESPACE>(bytes, n, doc_off, n_docs, cls_tbl, start_bits, drop_bits, fallbacks);
  else if (kind == PT_NOREGEX) emul_fast<PT_NOREGEX>(bytes, n, doc_off, n_docs, cls_tbl, start_bits, drop_bits, fallbacks);
  else if (kind == PT_BERT) emul_fast<PT_BERT>(bytes, n, doc_off, n_docs, cls_tbl, start_bits, drop_bits, fallbacks);
  else return 1;
  return 0;
}

extern "C" int b2t_emul_fast_planes(const uint8_t* bytes, uint64_t n, const uint64_t* doc_off, uint32_t n_docs, const uint32_t* cls_tbl,
                                    uint32_t* start_bits, uint32_t* planes_out) {
  emul_fast<PT_GPT2>(bytes, n, doc_off, n_docs, cls_tbl, start_bits, nullptr, nullptr, planes_out);
  return 0;
}

// Every code point >= 0x80 through classify_planes at every alignment of interest: wherever the boolean classifier
// claims a class (no `unc` bit), the claim must equal the class table.  cls8: 0x110000 bytes, one class per code point.
// Returns the number of wrong claims; *first_bad = first offending code point; *n_certain = code points with a claim.
template <int KIND>
static uint64_t check_claims(const uint8_t* cls8, uint32_t* first_bad, uint64_t* n_certain) {
  uint64_t bad = 0, certain = 0;
  for (uint32_t cp = 0x80; cp < 0x110000; ++cp) {
    if (cp >= 0xD800 && cp < 0xE000) continue;
    uint8_t enc[4]; int len;
    if (cp < 0x800) { enc[0] = 0xC0 | (cp >> 6); enc[1] = 0x80 | (cp & 63); len = 2; }
    else if (cp < 0x10000) { enc[0] = 0xE0 | (cp >> 12); enc[1] = 0x80 | ((cp >> 6) & 63); enc[2] = 0x80 | (cp & 63); len = 3; }
    else { enc[0] = 0xF0 | (cp >> 18); enc[1] = 0x80 | ((cp >> 12) & 63); enc[2] = 0x80 | ((cp >> 6) & 63); enc[3] = 0x80 | (cp & 63); len = 4; }
    bool any = false;
    for (int pos : {0, 7, 28, 29, 30, 31}) {
      uint8_t buf[32];
      std::memset(buf, 'a', 32);
      for (int k = 0; k < len && pos + k < 32; ++k) buf[pos + k] = enc[k];
      uint32_t w[8], b[8];
      std::memcpy(w, buf, 32);
      bitslice32(w, b);
      FastCls m = classify_planes<KIND>(b, 0xFFFFFFFFu);
      const uint32_t bit = 1u << pos;
      if (m.unc & bit) continue;
      any = true;
      const uint32_t claimed = (m.L & bit) ? CLS_L : (m.N & bit) ? CLS_N : (m.S & bit) ? CLS_S : CLS_O;
      if (claimed != cls8[cp]) { if (!bad) *first_bad = cp; ++bad; }
    }
    certain += any;
  }
  *n_certain = certain;
  return bad;
}
extern "C" uint64_t b2t_emul_check_claims(int kind, const uint8_t* cls8, uint32_t* first_bad, uint64_t* n_certain) {
  return kind == PT_WHITESPACE ? check_claims<PT_WHITESPACE>(cls8, first_bad, n_certain) : check_claims<PT_GPT2>(cls8, first_bad, n_certain);
}

// bitslice32 against the definition
extern "C" int b2t_emul_check_bitslice(uint64_t seed, int rounds) {
  uint64_t s = seed;
  for (int r = 0; r < rounds; ++r) {
    uint8_t buf[32];
    for (int i = 0; i < 32; ++i) { s = s * 6364136223846793005ull + 1442695040888963407ull; buf[i] = (uint8_t)(s >> 56); }
    uint32_t w[8], b[8];
    std::memcpy(w, buf, 32);
    bitslice32(w, b);
    for (int j = 0; j < 8; ++j) {
      uint32_t e = 0;
      for (int p = 0; p < 32; ++p) e |= (uint32_t)((buf[p] >> j) & 1u) << p;
      if (e != b[j]) return 1;
    }
  }
  return 0;
}
