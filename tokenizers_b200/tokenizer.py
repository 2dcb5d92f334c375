"""Host-side mirror of the reference's Python surface for the encode_batch path.

Mirrors `tokenizers.Tokenizer` (bindings/python/src/tokenizer.rs:510-1461 in the reference tree): `from_file`,
`from_str`, `encode`, `encode_batch`, `encode_batch_fast`, `token_to_id`, `id_to_token`, `get_vocab_size`, and
`tokenizers.Encoding` (bindings/python/src/encoding.rs:133-225): `ids`, `tokens`, `offsets`, `word_ids`, `type_ids`,
`attention_mask`, `special_tokens_mask`.  All compute happens in libb2t.so on the GPU; configurations outside the
hot path raise `UnsupportedConfig` (there is no CPU fallback).
"""
import ctypes, json
import numpy as np
from . import _lib
from ._lib import B2TError

LLAMA3_PATTERN = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*"
                  r"|\s*[\r\n]+|\s+(?!\S)|\s+")


class UnsupportedConfig(ValueError):
    """The tokenizer.json asks for something outside the accelerated path (the reference handles it on CPU)."""


def _pack(strings):
    bs = [s.encode("utf-8") for s in strings]
    off = np.zeros(len(bs) + 1, dtype=np.uint32)
    if bs:
        np.cumsum(np.fromiter(map(len, bs), dtype=np.int64, count=len(bs)), out=off[1:])
    return np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8).copy(), off


def parse_tokenizer_json(js):
    """tokenizer.json (tokenizer/serialization.rs:15-48 in the reference) -> engine configuration dict."""
    if js.get("normalizer") is not None:
        raise UnsupportedConfig("normalizers stay on the host and are not part of the accelerated path")
    if js.get("truncation") is not None or js.get("padding") is not None:
        raise UnsupportedConfig("truncation / padding are host post-processing and not supported here")
    pp = js.get("post_processor")
    if pp is not None and not (pp.get("type") == "ByteLevel" and not pp.get("trim_offsets", True)):
        raise UnsupportedConfig("post-processors other than a no-op ByteLevel(trim_offsets=False) are not supported")
    pt, m = js.get("pre_tokenizer"), js["model"]
    cfg = dict(add_prefix_space=0, ignore_merges=0, unk=None, prefix="##", max_chars=100, merges=[])
    if pt is None:
        raise UnsupportedConfig("a pre_tokenizer is required")
    if pt["type"] == "ByteLevel":
        cfg["pretok"] = _lib.PRETOK_BYTELEVEL if pt.get("use_regex", True) else _lib.PRETOK_BYTELEVEL_NOREGEX
        cfg["add_prefix_space"] = int(pt.get("add_prefix_space", True))
    elif pt["type"] == "Whitespace":
        cfg["pretok"] = _lib.PRETOK_WHITESPACE
    elif pt["type"] == "Sequence" and len(pt.get("pretokenizers", [])) == 2:
        a, b = pt["pretokenizers"]
        ok = (a.get("type") == "Split" and a.get("pattern", {}).get("Regex") == LLAMA3_PATTERN and a.get("behavior") == "Isolated"
              and not a.get("invert", False) and b.get("type") == "ByteLevel" and not b.get("use_regex", True)
              and not b.get("add_prefix_space", True))
        if not ok:
            raise UnsupportedConfig("only Sequence[Split(<tiktoken/Llama-3 pattern>, Isolated), ByteLevel(use_regex=False)] is supported")
        cfg["pretok"] = _lib.PRETOK_LLAMA3
    else:
        raise UnsupportedConfig(f"pre_tokenizer {pt['type']} is not on the accelerated path")
    if m["type"] == "BPE":
        cfg["model"] = _lib.MODEL_BPE
        for k in ("dropout", "unk_token", "continuing_subword_prefix", "end_of_word_suffix"):
            if m.get(k):
                raise UnsupportedConfig(f"BPE option {k} is not on the accelerated path")
        if m.get("byte_fallback") or m.get("fuse_unk"):
            raise UnsupportedConfig("BPE byte_fallback / fuse_unk are not on the accelerated path")
        cfg["ignore_merges"] = int(m.get("ignore_merges", False))
        cfg["merges"] = [tuple(x.split(" ")) if isinstance(x, str) else tuple(x) for x in m["merges"]]
    elif m["type"] == "WordPiece":
        cfg["model"] = _lib.MODEL_WORDPIECE
        cfg["unk"] = m["unk_token"]
        cfg["prefix"] = m["continuing_subword_prefix"]
        cfg["max_chars"] = m["max_input_chars_per_word"]
    else:
        raise UnsupportedConfig(f"model {m['type']} is not on the accelerated path")
    cfg["vocab"] = m["vocab"]
    cfg["added_tokens"] = [t["content"] for t in js.get("added_tokens", [])]
    return cfg


class Encoding:
    """One sequence of the batch CSR, with the attribute names of `tokenizers.Encoding`."""
    __slots__ = ("_tok", "ids", "_offsets", "_word_ids")

    def __init__(self, tok, ids, offsets, word_ids):
        self._tok, self.ids, self._offsets, self._word_ids = tok, ids, offsets, word_ids

    def __len__(self):
        return len(self.ids)

    @property
    def offsets(self):
        if self._offsets is None:
            raise ValueError("offsets were not requested (encode_batch_fast)")
        return [tuple(x) for x in self._offsets.tolist()]

    @property
    def word_ids(self):
        return None if self._word_ids is None else self._word_ids.tolist()

    words = word_ids

    @property
    def tokens(self):
        return [self._tok.id_to_token(i) for i in self.ids]

    @property
    def type_ids(self):
        return [0] * len(self.ids)

    @property
    def attention_mask(self):
        return [1] * len(self.ids)

    @property
    def special_tokens_mask(self):
        return [0] * len(self.ids)

    @property
    def sequence_ids(self):
        return [0] * len(self.ids)

    @property
    def n_sequences(self):
        return 1

    @property
    def overflowing(self):
        return []

    def __repr__(self):
        return f"Encoding(num_tokens={len(self.ids)}, attributes=[ids, type_ids, tokens, offsets, attention_mask, special_tokens_mask, overflowing])"


class BatchEncoding:
    """The whole batch as a CSR (numpy views copied out of the engine's pinned buffers)."""

    def __init__(self, ids, offsets, word_ids, row_ptr):
        self.ids, self.offsets, self.word_ids, self.row_ptr = ids, offsets, word_ids, row_ptr

    @property
    def n_tokens(self):
        return int(self.row_ptr[-1])


def _view(ptr, count, dtype):
    if not ptr or count == 0:
        return np.zeros(0, dtype=dtype)
    nbytes = count * np.dtype(dtype).itemsize
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(nbytes,)).view(dtype)


class Tokenizer:
    def __init__(self, tokenizer_json, device=-1):
        js = json.loads(tokenizer_json) if isinstance(tokenizer_json, (str, bytes)) else tokenizer_json
        cfg = parse_tokenizer_json(js)
        self._cfg = cfg
        self._vocab = cfg["vocab"]
        self._vocab_r = None
        self._added = cfg["added_tokens"]
        L = _lib.lib()
        toks = list(self._vocab.keys())
        vb, vo = _pack(toks)
        vi = np.fromiter((self._vocab[t] for t in toks), dtype=np.uint32, count=len(toks))
        mb, mo = _pack([s for ab in cfg["merges"] for s in ab])
        c = _lib.Config()
        c.struct_size = ctypes.sizeof(_lib.Config)
        c.model, c.pretok = cfg["model"], cfg["pretok"]
        c.add_prefix_space, c.ignore_merges = cfg["add_prefix_space"], cfg["ignore_merges"]
        c.n_vocab, c.vocab_bytes, c.vocab_off, c.vocab_ids = len(toks), vb.ctypes.data, vo.ctypes.data, vi.ctypes.data
        c.n_merges, c.merge_bytes, c.merge_off = len(cfg["merges"]), mb.ctypes.data, mo.ctypes.data
        c.unk_token = cfg["unk"].encode("utf-8") if cfg["unk"] is not None else None
        c.continuing_subword_prefix = cfg["prefix"].encode("utf-8")
        c.max_input_chars_per_word = cfg["max_chars"]
        c.device = device
        h = ctypes.c_void_p()
        rc = L.b2t_engine_create(ctypes.byref(c), ctypes.byref(h))
        if rc == _lib.B2T_ERR_UNSUPPORTED:
            raise UnsupportedConfig(L.b2t_last_error().decode())
        _lib.check(rc)
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            _lib.lib().b2t_engine_destroy(h)
            self._h = None

    # ---- construction (tokenizer/mod.rs:468-476)
    @staticmethod
    def from_str(s, device=-1):
        return Tokenizer(s, device)

    @staticmethod
    def from_file(path, device=-1):
        import gzip
        opener = gzip.open if str(path).endswith(".gz") else open
        with opener(path, "rb") as f:
            return Tokenizer(f.read().decode("utf-8"), device)

    # ---- vocabulary helpers
    def get_vocab_size(self, with_added_tokens=True):
        return len(self._vocab)

    def token_to_id(self, token):
        return self._vocab.get(token)

    def id_to_token(self, i):
        if self._vocab_r is None:
            self._vocab_r = {v: k for k, v in self._vocab.items()}
        return self._vocab_r.get(int(i))

    @property
    def handle(self):
        return self._h

    # ---- encode
    def _check_added(self, joined):
        for t in self._added:
            if t and t in joined:
                raise UnsupportedConfig(f"input contains the added token {t!r}: added-token extraction "
                                        "(added_vocabulary.rs:523-564) runs on the host in the reference and is not supported here yet")

    def encode_batch_csr(self, data, doc_off, offsets=True, word_ids=True, byte_offsets=False):
        """Packed batch in (np.uint8[N], np.uint64[n+1]) -> BatchEncoding.  Host buffers; copies happen inside."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        doc_off = np.ascontiguousarray(doc_off, dtype=np.uint64)
        n_docs = len(doc_off) - 1
        flags = (_lib.WANT_OFFSETS if offsets else 0) | (_lib.WANT_WORD_IDS if word_ids else 0) | (_lib.OFFSETS_BYTES if byte_offsets else 0)
        L = _lib.lib()
        res = ctypes.c_void_p()
        _lib.check(L.b2t_encode_batch(self._h, data.ctypes.data if data.size else None, doc_off.ctypes.data, n_docs, flags, ctypes.byref(res)))
        try:
            T = L.b2t_result_n_tokens(res)
            ids = _view(L.b2t_result_ids(res), T, np.uint32).copy()
            offs = _view(L.b2t_result_offsets(res), 2 * T, np.uint32).reshape(-1, 2).copy() if offsets else None
            wid = _view(L.b2t_result_word_ids(res), T, np.uint32).copy() if word_ids else None
            rp = _view(L.b2t_result_row_ptr(res), n_docs + 1, np.uint64).copy()
        finally:
            L.b2t_result_free(res)
        return BatchEncoding(ids, offs, wid, rp)

    def _encode_list(self, docs, offsets, word_ids):
        for d in docs:
            if not isinstance(d, str):
                raise UnsupportedConfig("only raw single sequences (str) are supported; pairs and pre-tokenized input are not")
        bs = [d.encode("utf-8") for d in docs]
        joined = b"".join(bs)
        if self._added:
            self._check_added(joined.decode("utf-8"))
        off = np.zeros(len(bs) + 1, dtype=np.uint64)
        if bs:
            np.cumsum(np.fromiter(map(len, bs), dtype=np.int64, count=len(bs)), out=off[1:])
        be = self.encode_batch_csr(np.frombuffer(joined, dtype=np.uint8), off, offsets, word_ids)
        rp = be.row_ptr
        out = []
        for i in range(len(docs)):
            a, b = int(rp[i]), int(rp[i + 1])
            out.append(Encoding(self, be.ids[a:b].tolist(), None if be.offsets is None else be.offsets[a:b],
                                None if be.word_ids is None else be.word_ids[a:b]))
        return out

    def encode_batch(self, input, is_pretokenized=False, add_special_tokens=True):
        """tokenizer.rs:1312-1340 -> TokenizerImpl::encode_batch_char_offsets (tokenizer/mod.rs:1360-1379)."""
        if is_pretokenized:
            raise UnsupportedConfig("is_pretokenized=True is not on the accelerated path")
        return self._encode_list(list(input), True, True)

    def encode_batch_fast(self, input, is_pretokenized=False, add_special_tokens=True):
        """tokenizer.rs:1433-1461 -> encode_batch_fast (tokenizer/mod.rs:1382-1401): ids only."""
        if is_pretokenized:
            raise UnsupportedConfig("is_pretokenized=True is not on the accelerated path")
        return self._encode_list(list(input), False, False)

    def encode(self, sequence, pair=None, is_pretokenized=False, add_special_tokens=True):
        if pair is not None or is_pretokenized:
            raise UnsupportedConfig("pairs / pre-tokenized input are not on the accelerated path")
        return self._encode_list([sequence], True, True)[0]

    def pre_tokenize_batch(self, docs):
        """PreTokenizer seam: per document the list of (start_byte, end_byte) of its splits."""
        bs = [d.encode("utf-8") for d in docs]
        off = np.zeros(len(bs) + 1, dtype=np.uint64)
        if bs:
            np.cumsum(np.fromiter(map(len, bs), dtype=np.int64, count=len(bs)), out=off[1:])
        data = np.frombuffer(b"".join(bs), dtype=np.uint8)
        L = _lib.lib()
        res = ctypes.c_void_p()
        _lib.check(L.b2t_pre_tokenize_batch(self._h, data.ctypes.data if data.size else None, off.ctypes.data, len(bs), ctypes.byref(res)))
        try:
            T = L.b2t_result_n_tokens(res)
            offs = _view(L.b2t_result_offsets(res), 2 * T, np.uint32).reshape(-1, 2).copy()
            rp = _view(L.b2t_result_row_ptr(res), len(bs) + 1, np.uint64).copy()
        finally:
            L.b2t_result_free(res)
        return [[tuple(x) for x in offs[int(rp[i]):int(rp[i + 1])].tolist()] for i in range(len(bs))]
