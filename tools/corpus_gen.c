// Synthetic corpus generator for the encode_batch benchmarks and parity tests.
//
// Test / bench infrastructure, not product code.  Deterministic: every document is generated from
// (seed, doc_index) alone, so any rank can generate any slice of a corpus without the rest.
//
// Corpus kinds (BASELINE.json `configs`):
//   1  plumbing : ASCII lines U[20,120] B, 5 000-word Zipf(1.0) lexicon, ".,!?'" punctuation
//   2  gpt2     : docs ~ lognormal(ln 400, 0.6) clipped [16, 8192] B, 200 k-word Zipf(1.1) lexicon
//                 (80 % ASCII, 20 % multi-byte), digit runs, punctuation + contractions, mixed whitespace
//   4  wordpiece: kind 2 lower-cased, ASCII + Latin-1 only, 0.5 % words > 100 chars, 1 % OOV-char words
//   5  skew     : doc lengths Zipf(1.2) over [8 B, 64 KB]; 0.1 % of docs hold one 4-64 KB letter / space run
//   6  gpt2 + special token: kind 2, every document ends with "<|endoftext|>" (the added-token extraction has work in every document)
//
// Build: gcc -O2 -shared -fPIC -o tools/libcorpus.so tools/corpus_gen.c -lm
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint64_t s;
} rng_t;

static inline uint64_t rng_next(rng_t *r) {  // splitmix64
  uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline double rng_unit(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint32_t rng_below(rng_t *r, uint32_t n) { return (uint32_t)(((rng_next(r) >> 32) * (uint64_t)n) >> 32); }

// ------------------------------------------------------------------ lexicon
typedef struct {
  uint32_t n_words;
  uint32_t *off;  // n_words + 1
  uint8_t *bytes;
  double *cdf;  // Zipf CDF
  uint8_t *ascii;  // 1 if word is pure a-z
} lexicon_t;

static int put_utf8(uint8_t *p, uint32_t c) {
  if (c < 0x80) { p[0] = (uint8_t)c; return 1; }
  if (c < 0x800) { p[0] = 0xC0 | (c >> 6); p[1] = 0x80 | (c & 63); return 2; }
  if (c < 0x10000) { p[0] = 0xE0 | (c >> 12); p[1] = 0x80 | ((c >> 6) & 63); p[2] = 0x80 | (c & 63); return 3; }
  p[0] = 0xF0 | (c >> 18); p[1] = 0x80 | ((c >> 12) & 63); p[2] = 0x80 | ((c >> 6) & 63); p[3] = 0x80 | (c & 63);
  return 4;
}

static const char *CONS[] = {"b", "c", "d", "f", "g", "h", "j", "k", "l", "m", "n", "p", "r", "s", "t", "v", "w",
                             "z", "th", "st", "ch", "sh", "tr", "pr", "nd", "ng", "qu", "x", "y", "ll"};
static const char *VOW[] = {"a", "e", "i", "o", "u", "ea", "ou", "ie", "oo", "ai"};

static int gen_ascii_word(rng_t *r, uint8_t *p, int target) {
  int n = 0;
  int cons = rng_below(r, 3) != 0;
  while (n < target) {
    const char *s = cons ? CONS[rng_below(r, 30)] : VOW[rng_below(r, 10)];
    for (; *s && n < target; ++s) p[n++] = (uint8_t)*s;
    cons = !cons;
  }
  return n;
}

static int gen_multibyte_word(rng_t *r, uint8_t *p, int latin_only) {
  int kind = latin_only ? 0 : (int)rng_below(r, 5);
  int n = 0;
  if (kind == 0) {  // Latin-1 accents mixed into an ASCII word
    static const uint32_t ACC[] = {0xE9, 0xE8, 0xE0, 0xFC, 0xF6, 0xE4, 0xF1, 0xE7, 0xEA, 0xF4, 0xED, 0xF3, 0xDF, 0xE5};
    int len = 3 + rng_below(r, 8), placed = 0;
    for (int i = 0; i < len; ++i) {
      if (rng_below(r, 4) == 0 || (i == len - 1 && !placed)) { n += put_utf8(p + n, ACC[rng_below(r, 14)]); placed = 1; }
      else p[n++] = 'a' + rng_below(r, 26);
    }
  } else if (kind == 1) {  // Cyrillic
    int len = 2 + rng_below(r, 9);
    for (int i = 0; i < len; ++i) n += put_utf8(p + n, 0x430 + rng_below(r, 32));
  } else if (kind == 2) {  // Greek
    int len = 2 + rng_below(r, 8);
    for (int i = 0; i < len; ++i) {
      uint32_t c = 0x3B1 + rng_below(r, 25);
      if (c == 0x3C2) c = 0x3C3;
      n += put_utf8(p + n, c);
    }
  } else if (kind == 3) {  // CJK
    int len = 1 + rng_below(r, 4);
    for (int i = 0; i < len; ++i) n += put_utf8(p + n, 0x4E00 + rng_below(r, 3000));
  } else {  // emoji
    int len = 1 + rng_below(r, 2);
    for (int i = 0; i < len; ++i) n += put_utf8(p + n, 0x1F600 + rng_below(r, 64));
  }
  return n;
}

static lexicon_t *lexicon_build(uint32_t n_words, double zipf_s, int multibyte_pct, int latin_only, uint64_t seed) {
  lexicon_t *lx = (lexicon_t *)calloc(1, sizeof(*lx));
  lx->n_words = n_words;
  lx->off = (uint32_t *)malloc((n_words + 1) * sizeof(uint32_t));
  lx->bytes = (uint8_t *)malloc((size_t)n_words * 48);
  lx->cdf = (double *)malloc(n_words * sizeof(double));
  lx->ascii = (uint8_t *)malloc(n_words);
  rng_t r = {seed ^ 0xA5A5A5A5DEADBEEFull};
  uint32_t pos = 0;
  double tot = 0;
  for (uint32_t i = 0; i < n_words; ++i) {
    lx->off[i] = pos;
    if ((int)rng_below(&r, 100) < multibyte_pct) {
      pos += gen_multibyte_word(&r, lx->bytes + pos, latin_only);
      lx->ascii[i] = 0;
    } else {
      // frequent words are short, rare ones long, like natural text
      int lo = i < 200 ? 2 : 2, hi = i < 200 ? 5 : 10;
      pos += gen_ascii_word(&r, lx->bytes + pos, lo + rng_below(&r, hi - lo + 1));
      lx->ascii[i] = 1;
    }
    tot += 1.0 / pow((double)(i + 1), zipf_s);
    lx->cdf[i] = tot;
  }
  lx->off[n_words] = pos;
  for (uint32_t i = 0; i < n_words; ++i) lx->cdf[i] /= tot;
  return lx;
}

static void lexicon_free(lexicon_t *lx) {
  free(lx->off); free(lx->bytes); free(lx->cdf); free(lx->ascii); free(lx);
}

static uint32_t lexicon_sample(const lexicon_t *lx, rng_t *r) {
  double u = rng_unit(r);
  uint32_t lo = 0, hi = lx->n_words - 1;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (lx->cdf[mid] < u) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------ documents
static const char *CONTR[] = {"'s", "'t", "'re", "'ve", "'m", "'ll", "'d", "'S", "'T", "'RE", "'VE", "'M", "'LL", "'D"};
static const char PUNCT[] = ".,!?;:-()\"'/&%$#@*";

typedef struct {
  int kind;
  const lexicon_t *lx;
} gen_t;

static uint32_t emit_separator(rng_t *r, uint8_t *p, int kind) {
  uint32_t u = rng_below(r, 100);
  if (kind == 1) { p[0] = ' '; return 1; }
  if (u < 90) { p[0] = ' '; return 1; }
  if (u < 94) { p[0] = '\n'; return 1; }
  if (u < 96) { p[0] = ' '; p[1] = ' '; return 2; }
  if (u < 98) { p[0] = '\n'; p[1] = '\n'; return 2; }
  if (u < 99) { p[0] = '\t'; return 1; }
  if (kind == 4) { p[0] = '\r'; p[1] = '\n'; return 2; }
  return rng_below(r, 2) ? put_utf8(p, 0xA0) : put_utf8(p, 0x3000);
}

// Generates one document of roughly `target` bytes (never more than target + 160) into p.
static uint32_t gen_doc(const gen_t *g, rng_t *r, uint8_t *p, uint32_t target) {
  uint32_t n = 0;
  const lexicon_t *lx = g->lx;
  while (n < target) {
    uint32_t u = rng_below(r, 1000);
    if (g->kind == 1) {
      uint32_t w = lexicon_sample(lx, r);
      uint32_t len = lx->off[w + 1] - lx->off[w];
      memcpy(p + n, lx->bytes + lx->off[w], len); n += len;
      if (rng_below(r, 10) == 0) p[n++] = ".,!?'"[rng_below(r, 5)];
    } else if (u < 880) {  // word
      uint32_t w = lexicon_sample(lx, r);
      uint32_t len = lx->off[w + 1] - lx->off[w];
      memcpy(p + n, lx->bytes + lx->off[w], len);
      if (g->kind != 4 && lx->ascii[w] && rng_below(r, 100) < 15) p[n] -= 32;  // capitalise
      n += len;
      if (g->kind == 4) {
        uint32_t v = rng_below(r, 1000);
        if (v < 5) {  // word longer than max_input_chars_per_word
          uint32_t extra = 101 + rng_below(r, 40);
          for (uint32_t i = 0; i < extra; ++i) p[n++] = 'a' + rng_below(r, 26);
        } else if (v < 15) {  // OOV char inside the word (never in the training alphabet)
          n += put_utf8(p + n, 0x0F00 + rng_below(r, 40));
        }
      }
      if (rng_below(r, 100) < 3) {  // contraction
        const char *c = CONTR[rng_below(r, g->kind == 4 ? 7 : 14)];
        for (; *c; ++c) p[n++] = (uint8_t)*c;
      }
    } else if (u < 940) {  // digit run
      uint32_t len = 1 + rng_below(r, 8);
      for (uint32_t i = 0; i < len; ++i) p[n++] = '0' + rng_below(r, 10);
    } else {  // punctuation run
      uint32_t len = 1 + (rng_below(r, 4) == 0 ? rng_below(r, 3) : 0);
      for (uint32_t i = 0; i < len; ++i) p[n++] = PUNCT[rng_below(r, sizeof(PUNCT) - 1)];
    }
    if (n < target) n += emit_separator(r, p + n, g->kind);
  }
  return n;
}

static uint32_t doc_target_len(int kind, rng_t *r) {
  if (kind == 1) return 20 + rng_below(r, 101);
  if (kind == 5) {
    // Zipf(1.2) over [8, 65536]: inverse-CDF of the continuous power law x^-1.2
    double u = rng_unit(r), a = 1.2, lo = 8.0, hi = 65536.0;
    double x = pow(pow(lo, 1 - a) + u * (pow(hi, 1 - a) - pow(lo, 1 - a)), 1.0 / (1 - a));
    return (uint32_t)x;
  }
  double u1 = rng_unit(r), u2 = rng_unit(r);
  if (u1 < 1e-300) u1 = 1e-300;
  double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
  double len = exp(log(400.0) + 0.6 * z);
  if (len < 16) len = 16;
  if (len > 8192) len = 8192;
  return (uint32_t)len;
}

typedef struct {
  gen_t g;
  lexicon_t *lx;
  uint64_t seed;
} corpus_t;

// Opaque handle: building the lexicon costs ~0.1 s, so keep it across calls.
void *b2t_corpus_open(int kind, uint64_t seed) {
  corpus_t *c = (corpus_t *)calloc(1, sizeof(*c));
  c->seed = seed;
  c->g.kind = kind;
  if (kind == 1) c->lx = lexicon_build(5000, 1.0, 0, 0, 1);
  else if (kind == 4) c->lx = lexicon_build(200000, 1.1, 12, 1, 4);
  else c->lx = lexicon_build(200000, 1.1, 20, 0, 2);
  c->g.lx = c->lx;
  return c;
}

void b2t_corpus_close(void *h) {
  corpus_t *c = (corpus_t *)h;
  lexicon_free(c->lx);
  free(c);
}

// Generate documents [first_doc, first_doc + n_docs).  `out` must hold `cap` bytes; doc_off gets n_docs + 1 entries
// (relative to out).  Stops early (returning the number of docs written) if the next doc might not fit.
uint64_t b2t_corpus_generate(void *h, uint64_t first_doc, uint64_t n_docs, uint8_t *out, uint64_t cap, uint64_t *doc_off) {
  corpus_t *c = (corpus_t *)h;
  uint64_t pos = 0, d;
  doc_off[0] = 0;
  for (d = 0; d < n_docs; ++d) {
    rng_t r = {c->seed * 0x9E3779B97F4A7C15ull + (first_doc + d) * 0xD1B54A32D192ED03ull + 0x1234567ull};
    rng_next(&r);
    uint32_t target = doc_target_len(c->g.kind, &r);
    if (pos + (uint64_t)target + 140000 > cap) break;
    uint32_t n;
    if (c->g.kind == 5 && rng_below(&r, 1000) == 0) {
      // one giant pre-token inside an otherwise normal doc
      n = gen_doc(&c->g, &r, out + pos, target / 2);
      out[pos + n++] = ' ';
      uint32_t run = 4096 + rng_below(&r, 61440);
      if (rng_below(&r, 2)) { for (uint32_t i = 0; i < run; ++i) out[pos + n++] = 'a' + rng_below(&r, 26); }
      else { for (uint32_t i = 0; i < run; ++i) out[pos + n++] = ' '; }
      out[pos + n++] = ' ';
      n += gen_doc(&c->g, &r, out + pos + n, target / 2 + 1);
    } else {
      n = gen_doc(&c->g, &r, out + pos, target);
    }
    if (c->g.kind == 6) { memcpy(out + pos + n, "<|endoftext|>", 13); n += 13; }
    pos += n;
    doc_off[d + 1] = pos;
  }
  return d;
}
