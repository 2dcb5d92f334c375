"""CPU: the bit-sliced fast path of the pre-tokenization scan (tokenizers_b200/csrc/pretok_fast.cuh) executed on the host
by tests/native/pretok_emul.cpp: the bit-plane transposition against its definition, every claim of the boolean
non-ASCII classifier against the class tables (all code points), and the composed chunk pipeline fuzzed against the
oracle's regex restatement and against the window code it replaces."""
import ctypes
import numpy as np
import pytest
import helpers, fuzzgen, corpus
from oracle import oracle as orc
from test_pretok_logic_cpu import _emul, _pack2, _expected, CFG


def _lib():
    L = _emul()
    L.b2t_emul_pretok_fast.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint32,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.b2t_emul_check_claims.restype = ctypes.c_uint64
    L.b2t_emul_check_claims.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.b2t_emul_check_bitslice.argtypes = [ctypes.c_uint64, ctypes.c_int]
    return L


def test_bitslice_is_a_transposition():
    assert _lib().b2t_emul_check_bitslice(12345, 20000) == 0


@pytest.mark.parametrize("scheme,kind", [("onig", 0), ("rust", 2)])
def test_boolean_classifier_never_contradicts_the_table(scheme, kind):
    L = _lib()
    t = np.ascontiguousarray(orc.class_table(scheme), dtype=np.uint8)
    bad = ctypes.c_uint32(0); certain = ctypes.c_uint64(0)
    n_bad = L.b2t_emul_check_claims(kind, t.ctypes.data, ctypes.byref(bad), ctypes.byref(certain))
    assert n_bad == 0, f"{scheme}: {n_bad} wrong claims, first U+{bad.value:04X}"
    # the fast classifier must actually cover the big blocks (Latin, Greek, Cyrillic, CJK, Hangul, emoji)
    assert certain.value > 60000, certain.value


def _straddle_docs(seed, count):
    """Documents dense in multi-byte whitespace, apostrophes and short runs, so that characters and contractions straddle
    32-byte chunk ends in every alignment."""
    rng = np.random.default_rng(seed)
    pieces = [" ", "\u3000", "\u00a0", "\u2003", "\u2028", "  ", "\n", "a", "b's", "'ll", "'re", "'t", "'", "x'", "1",
              "\u00e9", "\u044f", "\u4e2d", "\U0001F600", "\u00d7", "!", "don't", "we've", "I'M", "'S", "\t", "ab",
              "\u0085", "\u1680", "\u0663", "\u2167", "\u00aa", "\u03a9", "\u03f6", "\u4dc5", "\ud55c", "\u2014", "\u3042",
              "\U0001F100", "\u017f", "'\u017f"]
    docs = []
    for _ in range(count):
        k = int(rng.integers(1, 40))
        docs.append("".join(pieces[int(i)] for i in rng.integers(0, len(pieces), k)))
    return docs


def _bert_json():
    import json
    js = json.loads(helpers.asset_json("wordpiece"))
    js["pre_tokenizer"] = {"type": "BertPreTokenizer"}
    return json.dumps(js)


@pytest.mark.parametrize("name", ["gpt2_style", "wordpiece", "bert_pretok"])
def test_fast_path_matches_oracle(name):
    L = _lib()
    kind = 4 if name == "bert_pretok" else CFG[name]
    o = orc.Oracle(_bert_json() if kind == 4 else helpers.asset_json(name))
    tbl = _pack2(orc.class_table("bert" if kind == 4 else ("rust" if kind == 2 else "onig")))
    batches = [fuzzgen.rand_docs(700 + s, 600, max_len=60 if s % 3 else 400) for s in range(6)]
    batches += [_straddle_docs(40 + s, 500) for s in range(4)]
    for k in (1, 2, 4, 5):
        data, off = corpus.generate(k, 50 + k, 0, 150)
        batches.append(corpus.to_strings(data, off))
    total_fb = 0
    for docs in batches:
        data, off = helpers.pack_docs(docs)
        n = int(off[-1])
        buf = np.concatenate([data, np.zeros(64, dtype=np.uint8)])
        st = np.zeros(n // 32 + 2, dtype=np.uint32); dr = np.zeros(n // 32 + 2, dtype=np.uint32)
        fb = ctypes.c_uint64(0)
        assert L.b2t_emul_pretok_fast(kind, buf.ctypes.data, n, off.ctypes.data, len(docs), tbl.ctypes.data, st.ctypes.data,
                                      dr.ctypes.data, ctypes.byref(fb)) == 0
        total_fb += fb.value
        bits = np.unpackbits(st.view(np.uint8), bitorder="little")[:n]
        dbits = np.unpackbits(dr.view(np.uint8), bitorder="little")[:n]
        exp, expd = _expected(o, kind, docs, off, n)
        bad = np.nonzero((bits != exp) | (dbits != expd))[0]
        if len(bad):
            di = int(np.searchsorted(off, bad[0], side="right") - 1)
            raise AssertionError(f"{name}: boundary mismatch at byte {bad[0] - int(off[di])} of doc {docs[di]!r}")
        if kind == 4:
            continue   # (the Bert pre-tokenizer has no window form)
        # and bit for bit the window code
        st2 = np.zeros_like(st); dr2 = np.zeros_like(dr)
        L.b2t_emul_pretok(kind, buf.ctypes.data, n, off.ctypes.data, len(docs), tbl.ctypes.data, st2.ctypes.data, dr2.ctypes.data)
        assert np.array_equal(st[: n // 32 + 1], st2[: n // 32 + 1]) and np.array_equal(dr[: n // 32 + 1], dr2[: n // 32 + 1])
    if kind == 0:
        assert total_fb > 0   # the straddling multi-byte space case was exercised
