/* b2t.h -- C ABI of the B200 batched tokenization engine (libb2t.so).
 *
 * Drop-in boundary for ONE path of huggingface/tokenizers: `Tokenizer::encode_batch` for ByteLevel-BPE
 * (GPT-2 / Llama-3 style) and Whitespace + WordPiece.  Each entry point names the reference interface it
 * replaces (paths relative to /root/reference/tokenizers/src unless noted).  A Rust host would bind this file
 * with an `extern "C"` block; the Python shim in tokenizers_b200/ binds it with ctypes.
 *
 * Conventions: plain pointers and sizes only; every function returns a b2t_status value (0 = ok) unless noted; the
 * message of the last error on the calling thread is available from b2t_last_error().  No CPU fallback exists:
 * configurations outside the supported set fail with B2T_ERR_UNSUPPORTED, and every encode entry point needs a
 * CUDA device (sm_100a).
 */
#ifndef B2T_H_
#define B2T_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2t_engine b2t_engine;
typedef struct b2t_result b2t_result;

typedef enum {
  B2T_OK = 0,
  B2T_ERR_INVALID = 1,     /* bad argument */
  B2T_ERR_UNSUPPORTED = 2, /* configuration outside the hot path (the reference would handle it on CPU) */
  B2T_ERR_CUDA = 3,        /* CUDA runtime error, message has the details */
  B2T_ERR_VOCAB = 4,       /* merge token out of vocabulary / missing [UNK] (models/bpe/mod.rs:12-36, wordpiece/mod.rs:17-22) */
  B2T_ERR_TOO_LARGE = 5    /* batch exceeds the per-call device limits */
} b2t_status;

/* models::ModelWrapper variants on the path (models/mod.rs:60-68) */
typedef enum { B2T_MODEL_BPE = 0, B2T_MODEL_WORDPIECE = 1 } b2t_model_kind;

/* pre_tokenizers::PreTokenizerWrapper configurations on the path (pre_tokenizers/mod.rs:28-62) */
typedef enum {
  B2T_PRETOK_BYTELEVEL = 0,         /* ByteLevel{use_regex=true}: GPT-2 pattern, byte_level.rs:43-46,119-148 */
  B2T_PRETOK_LLAMA3 = 1,            /* Sequence[Split(tiktoken pattern, Isolated), ByteLevel{use_regex=false}] */
  B2T_PRETOK_WHITESPACE = 2,        /* Whitespace: whitespace.rs:20-29 */
  B2T_PRETOK_BYTELEVEL_NOREGEX = 3, /* ByteLevel{use_regex=false}: the whole sequence is one pre-token */
  B2T_PRETOK_BERT = 4               /* BertPreTokenizer: pre_tokenizers/bert.rs:5-19 (whitespace removed, punctuation isolated) */
} b2t_pretok_kind;

/* normalizers::BertNormalizer (normalizers/bert.rs:52-136), in front of a WordPiece pipeline: b2t_config.bert_normalizer =
 * B2T_NORM_BERT | the enabled steps (strip_accents: None resolves to lowercase, bert.rs:128).  The normalizer runs on the
 * device; offsets refer to the ORIGINAL text through the alignments (tokenizer/normalizer.rs:317-428).  0 = no normalizer. */
enum { B2T_NORM_BERT = 0x100, B2T_NORM_CLEAN_TEXT = 1, B2T_NORM_CHINESE_CHARS = 2, B2T_NORM_STRIP_ACCENTS = 4, B2T_NORM_LOWERCASE = 8 };

/* Engine configuration = what `TokenizerBuilder` (tokenizer/mod.rs:315-437) receives for this path:
 * BPE::builder().vocab_and_merges(..).ignore_merges(..) (models/bpe/model.rs:36-210) or
 * WordPiece::builder().vocab(..).unk_token(..).continuing_subword_prefix(..).max_input_chars_per_word(..)
 * (models/wordpiece/mod.rs:40-120), plus the pre-tokenizer flags. Strings are UTF-8, exactly as in tokenizer.json. */
typedef struct {
  uint32_t struct_size; /* sizeof(b2t_config), for ABI evolution */
  int32_t model;        /* b2t_model_kind */
  int32_t pretok;       /* b2t_pretok_kind */
  int32_t add_prefix_space; /* ByteLevel.add_prefix_space (byte_level.rs:57-68) */
  int32_t ignore_merges;    /* BPE.ignore_merges (models/bpe/model.rs:558-567) */
  /* vocabulary: n_vocab token strings, packed back to back; token i = vocab_bytes[vocab_off[i] .. vocab_off[i+1]) */
  uint32_t n_vocab;
  const uint8_t* vocab_bytes;
  const uint32_t* vocab_off; /* n_vocab + 1 */
  const uint32_t* vocab_ids; /* n_vocab */
  /* BPE merges in rank order: merge i = (string 2i, string 2i+1) of the packed list (models/bpe/model.rs:252-275) */
  uint32_t n_merges;
  const uint8_t* merge_bytes;
  const uint32_t* merge_off; /* 2 * n_merges + 1 */
  /* WordPiece */
  const char* unk_token;                 /* NUL-terminated, may be NULL for BPE */
  const char* continuing_subword_prefix; /* NUL-terminated, e.g. "##" */
  uint32_t max_input_chars_per_word;     /* 100 in bert */
  int32_t device;                        /* CUDA device ordinal, -1 = current device */
  int32_t bert_normalizer;               /* B2T_NORM_* flags, 0 = none */
} b2t_config;

/* encode flags */
enum {
  B2T_WANT_OFFSETS = 1u,   /* produce (start, end) per token */
  B2T_WANT_WORD_IDS = 2u,  /* produce the pre-token ordinal per token (Encoding.words, tokenizer/pre_tokenizer.rs:252-256) */
  B2T_OFFSETS_BYTES = 4u,  /* OffsetType::Byte (Rust encode_batch) instead of OffsetType::Char (encode_batch_char_offsets,
                              what the Python binding always uses: bindings/python/src/tokenizer.rs:1332) */
  B2T_NO_ADDED_TOKENS = 8u,   /* skip the added-token extraction of an engine that has added tokens (the caller knows the text
                                 holds none, or has split it already) */
  B2T_FLAG_ADDED_IDS = 16u    /* mark the tokens that come from the added vocabulary with bit 31 of their id */
};

/* AddedToken properties (tokenizer/added_vocabulary.rs:11-76) */
enum { B2T_ADDED_SINGLE_WORD = 1u, B2T_ADDED_LSTRIP = 2u, B2T_ADDED_RSTRIP = 4u, B2T_ADDED_NORMALIZED = 8u };

/* Replaces TokenizerBuilder::build for the path.  The tables are uploaded to the device once; the engine is
 * immutable afterwards and may be used from several host threads: every host-buffer call (b2t_encode_batch,
 * b2t_encode_batch_dense, b2t_pre_tokenize_batch) runs on device workspaces and streams of its own (up to four calls in
 * flight, further ones wait), so their copies and kernels overlap; the device-resident entry points share one workspace
 * (their result lives in it) and serialise.  With profiling on (b2t_engine_set_profiling) calls are meant to be made one at a time. */
int b2t_engine_create(const b2t_config* cfg, b2t_engine** out);
void b2t_engine_destroy(b2t_engine* e);

/* Replaces AddedVocabulary::add_tokens / refresh_added_tokens (tokenizer/added_vocabulary.rs:270-420) for the path: the
 * tokens the encode entry points extract from the text BEFORE pre-tokenization, like extract_and_normalize without a
 * normalizer (added_vocabulary.rs:523-564: non-normalized tokens first, then the normalized ones on the remaining pieces;
 * leftmost-longest, single_word / lstrip / rstrip).  Token i = bytes[off[i] .. off[i+1]), ids[i] < 2^20, flags[i] = OR of
 * B2T_ADDED_*.  The extraction runs on the device (no re-packing: span boundaries become hard boundaries of the scan, a
 * span becomes one pre-token that carries the token's id).  Not combinable with add_prefix_space (B2T_ERR_UNSUPPORTED); a
 * batch whose spans do not fit the device limits (a span over 256 bytes after lstrip / rstrip, more spans than one per
 * 16 input bytes, the reference's overlapping-span corner case) fails with B2T_ERR_UNSUPPORTED -- the host can then split
 * the text itself and pass B2T_NO_ADDED_TOKENS.  n_tokens = 0 clears the set.  Not to be called concurrently with encodes. */
int b2t_engine_set_added_tokens(b2t_engine* e, uint32_t n_tokens, const uint8_t* bytes, const uint32_t* off,
                                const uint32_t* ids, const uint8_t* flags);

/* Replaces TokenizerImpl::encode_batch / encode_batch_char_offsets / encode_batch_fast (tokenizer/mod.rs:1337-1401)
 * for raw (not pre-tokenized) single sequences with add_special_tokens=false.  HOST buffers: `bytes` holds the
 * documents back to back, document d = bytes[doc_off[d] .. doc_off[d+1]); doc_off[0] must be 0 and doc_off must be
 * non-decreasing (checked: B2T_ERR_INVALID otherwise -- nothing unordered reaches a kernel).  The call copies the
 * input to the device in chunks, runs the kernels and copies the token CSR back into pinned host memory owned by
 * the result.  flags = 0 is the encode_batch_fast analogue (ids only). */
int b2t_encode_batch(b2t_engine* e, const uint8_t* bytes, const uint64_t* doc_off, uint32_t n_docs, uint32_t flags,
                     b2t_result** out);

/* Same, with the packed batch already in device memory (d_bytes 16-byte aligned, n_bytes = doc_off[n_docs] < 2^31) and the
 * result left in device memory (owned by the engine, valid until the next call on this engine or b2t_result_free).
 * `stream` is a cudaStream_t (NULL = the engine's own stream); the call is asynchronous with respect to the host
 * except for one small device-to-host read of the token count.  This is the entry point the roofline numbers use. */
int b2t_encode_batch_device(b2t_engine* e, const uint8_t* d_bytes, uint64_t n_bytes, const uint64_t* d_doc_off,
                            uint32_t n_docs, uint32_t flags, void* stream, b2t_result** out);

/* The same call in two halves, for callers that place the result themselves -- the multi-GPU path: the reference fans a
 * batch out over rayon threads (tokenizer/mod.rs:1345-1348) and collects the encodings in input order; here every rank
 * (one process per GPU) encodes its contiguous shard of the batch and writes its part of the token CSR straight at its
 * displacement of the buffer that the all-gather then completes (tokenizers_b200/parallel.py encode_batch_sharded).
 *   begin : everything up to the token counts; *n_tokens = tokens of this shard (one small device-to-host read).
 *   finish: writes ids[0..n_tokens), offsets[0..2 n_tokens) (if requested at begin), word_ids (if requested) and
 *           row_ptr[0..n_docs] = token_base + the shard's row_ptr to the DEVICE pointers given (already displaced by
 *           the caller); asynchronous on `stream`.  No other call on this engine between begin and finish. */
int b2t_encode_batch_device_begin(b2t_engine* e, const uint8_t* d_bytes, uint64_t n_bytes, const uint64_t* d_doc_off,
                                  uint32_t n_docs, uint32_t flags, void* stream, uint64_t* n_tokens);
int b2t_encode_batch_device_finish(b2t_engine* e, uint32_t* d_ids, uint32_t* d_offsets, uint32_t* d_word_ids,
                                   uint64_t* d_row_ptr, uint64_t token_base, void* stream);

/* Dense mode: the steps the reference runs AFTER the path for a batch of single sequences -- truncation
 * (utils/truncation.rs:70-166, kept part of Encoding::truncate, tokenizer/encoding.rs:307-388), the special-token template
 * `pre $A post` (processors/template.rs:646-, BertProcessing / RobertaProcessing single form) and padding
 * (utils/padding.rs:50-81), as TokenizerImpl::post_process orders them (tokenizer/mod.rs:1265-1317) -- done on the device:
 * the result is a dense [n_docs, L] tensor of ids (+ attention mask + row lengths) instead of the token CSR, which never
 * leaves the device.  Overflowing parts (stride) are not part of a dense batch; offsets are not produced. */
typedef struct {
  uint32_t struct_size;        /* sizeof(b2t_dense_spec) */
  uint32_t length;             /* PaddingStrategy::Fixed(length); 0 = BatchLongest (needs the batch in one device pass: < 2^31 bytes) */
  uint32_t pad_to_multiple_of; /* PaddingParams.pad_to_multiple_of, 0 = none */
  uint32_t max_length;         /* TruncationParams.max_length, special tokens included; 0 = no truncation */
  uint32_t pad_id;             /* PaddingParams.pad_id */
  int32_t truncate_left;       /* TruncationDirection::Left: keep the LAST tokens */
  int32_t pad_left;            /* PaddingDirection::Left */
  uint32_t n_pre, n_post;      /* special tokens of the single-sequence template before / after the sequence (<= 8 each) */
  const uint32_t* pre_ids;
  const uint32_t* post_ids;
  uint32_t want_mask;          /* also return the attention mask (u8 per position); row lengths always come back */
} b2t_dense_spec;

/* HOST buffers in (as b2t_encode_batch), pinned host rows out.  A row that does not fit L (Fixed length without a
 * sufficient truncation; the reference returns a longer row there) fails the batch with B2T_ERR_INVALID. */
int b2t_encode_batch_dense(b2t_engine* e, const uint8_t* bytes, const uint64_t* doc_off, uint32_t n_docs,
                           const b2t_dense_spec* spec, b2t_result** out);
/* Device buffers in (as b2t_encode_batch_device), device rows out (owned by the engine until the next call). */
int b2t_encode_batch_dense_device(b2t_engine* e, const uint8_t* d_bytes, uint64_t n_bytes, const uint64_t* d_doc_off,
                                  uint32_t n_docs, const b2t_dense_spec* spec, void* stream, b2t_result** out);
/* Dense results: row d = ids[d * L .. (d + 1) * L); row_lengths[d] = tokens of row d that are not padding. */
uint32_t b2t_result_dense_length(const b2t_result* r);        /* L */
const uint32_t* b2t_result_dense_ids(const b2t_result* r);    /* n_docs * L */
const uint8_t* b2t_result_attention_mask(const b2t_result* r); /* n_docs * L, or NULL */
const uint32_t* b2t_result_row_lengths(const b2t_result* r);  /* n_docs */

/* Replaces PreTokenizer::pre_tokenize (tokenizer/mod.rs:65-67) for a batch: the splits of every document as
 * (start, end) BYTE offsets into the document (offsets[2k], offsets[2k+1]); row_ptr delimits documents.  ids and
 * word_ids are absent.  Host buffers in, pinned host buffers out.  (With add_prefix_space the split that contains the
 * inserted space starts at the first byte of the document.) */
int b2t_pre_tokenize_batch(b2t_engine* e, const uint8_t* bytes, const uint64_t* doc_off, uint32_t n_docs,
                           b2t_result** out);

/* Result accessors (Encoding fields of tokenizer/encoding.rs:11-31 as one CSR over the batch).  Pointers are host
 * pointers for b2t_encode_batch / b2t_pre_tokenize_batch and device pointers for b2t_encode_batch_device. */
uint64_t b2t_result_n_tokens(const b2t_result* r);
uint32_t b2t_result_n_docs(const b2t_result* r);
int b2t_result_on_device(const b2t_result* r);
const uint32_t* b2t_result_ids(const b2t_result* r);      /* n_tokens */
const uint32_t* b2t_result_offsets(const b2t_result* r);  /* 2 * n_tokens, or NULL */
const uint32_t* b2t_result_word_ids(const b2t_result* r); /* n_tokens, or NULL */
const uint64_t* b2t_result_row_ptr(const b2t_result* r);  /* n_docs + 1 */
/* Host results return their pinned buffers to the engine's pool: free every result BEFORE b2t_engine_destroy. */
void b2t_result_free(b2t_result* r);

/* Pinned host memory for callers that want b2t_encode_batch to copy straight from their buffer. */
int b2t_host_alloc(size_t bytes, void** out);
void b2t_host_free(void* p);

/* Per-kernel device timings of the last b2t_encode_batch_device call when profiling is on (CUDA events on the
 * launch stream).  names/ms arrays of capacity cap; returns the number of kernels launched by that call. */
int b2t_engine_set_profiling(b2t_engine* e, int on);
int b2t_engine_last_kernels(const b2t_engine* e, const char** names, float* ms, int cap);

/* Unicode class tables the scan kernels use, one byte per code point (0x110000 entries):
 * scheme 0 (Oniguruma: ByteLevel / Split): 1 = \p{L}, 2 = \p{N}, 3 = \s, 0 = other;
 * scheme 1 (Rust regex: Whitespace):       1 = \w, 3 = \s, 0 = other;
 * scheme 2 (BertPreTokenizer):              1 = word character, 3 = whitespace (removed), 0 = punctuation (isolated).
 * Host only, no device needed. */
int b2t_unicode_class_table(int scheme, uint8_t* out);

/* The BertNormalizer table the device kernels use, for inspection (host only, no device needed): the UTF-8 image of every
 * code point under `flags` (B2T_NORM_* steps), images back to back in `pool` (capacity `cap`, 5 MB is enough), code point c =
 * pool[off[c] .. off[c+1]) with off holding 0x110001 entries; an empty image = the character is dropped. */
int b2t_bert_normalizer_images(int32_t flags, uint8_t* pool, size_t cap, uint32_t* off);

/* Thread-local message of the last failing call. */
const char* b2t_last_error(void);
/* Library version string. */
const char* b2t_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B2T_H_ */
