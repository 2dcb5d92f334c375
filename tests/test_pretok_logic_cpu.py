"""CPU: the device pre-tokenization mask logic (tokenizers_b200/csrc/pretok_logic.cuh) executed on the host by
tests/native/pretok_emul.cpp, fuzzed against the oracle's regex restatement."""
import ctypes, os, subprocess
import numpy as np
import pytest
import helpers, fuzzgen, corpus
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "pretok_emul.cpp")
SO = os.path.join(HERE, "native", "libpretok_emul.so")
HDR = os.path.join(helpers.ROOT, "tokenizers_b200", "csrc", "pretok_logic.cuh")


def _emul():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC])
    L = ctypes.CDLL(SO)
    L.b2t_emul_pretok.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint32,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return L


def _pack2(t):
    t = t.astype(np.uint32).reshape(-1, 16)
    return np.bitwise_or.reduce(t << (np.arange(16, dtype=np.uint32) * 2), axis=1).astype(np.uint32)


def _expected(o, kind, docs, off, n):
    exp = np.zeros(n + 1, dtype=np.uint8); expd = np.zeros(n + 1, dtype=np.uint8)
    for i, d in enumerate(docs):
        base, end = int(off[i]), int(off[i + 1])
        sp = o.pre_tokenize(d)
        if kind in (2, 4):   # Whitespace / BertPreTokenizer: the gaps between the splits are removed whitespace
            pos = 0
            for a, b in sp:
                if a > pos:
                    exp[base + pos] = 1; expd[base + pos] = 1
                exp[base + a] = 1; pos = b
            if pos < end - base:
                exp[base + pos] = 1; expd[base + pos] = 1
        else:
            for a, b in sp:
                exp[base + a] = 1
    return exp[:n], expd[:n]


CFG = {"gpt2_style": 0, "llama3_style": 1, "wordpiece": 2}


@pytest.mark.parametrize("name", list(CFG))
def test_mask_logic_matches_oracle(name):
    L = _emul()
    kind = CFG[name]
    o = orc.Oracle(helpers.asset_json(name))
    tbl = _pack2(orc.class_table("rust" if kind == 2 else "onig"))
    batches = [fuzzgen.rand_docs(200 + s, 600, max_len=60 if s % 3 else 400) for s in range(6)]
    for k in (1, 2, 4, 5):
        data, off = corpus.generate(k, 30 + k, 0, 120)
        batches.append(corpus.to_strings(data, off))
    for docs in batches:
        data, off = helpers.pack_docs(docs)
        n = int(off[-1])
        buf = np.concatenate([data, np.zeros(64, dtype=np.uint8)])
        st = np.zeros(n // 32 + 2, dtype=np.uint32); dr = np.zeros(n // 32 + 2, dtype=np.uint32)
        L.b2t_emul_pretok(kind, buf.ctypes.data, n, off.ctypes.data, len(docs), tbl.ctypes.data, st.ctypes.data, dr.ctypes.data)
        bits = np.unpackbits(st.view(np.uint8), bitorder="little")[:n]
        dbits = np.unpackbits(dr.view(np.uint8), bitorder="little")[:n]
        exp, expd = _expected(o, kind, docs, off, n)
        bad = np.nonzero((bits != exp) | (dbits != expd))[0]
        if len(bad):
            di = int(np.searchsorted(off, bad[0], side="right") - 1)
            raise AssertionError(f"{name}: boundary mismatch at byte {bad[0] - int(off[di])} of doc {docs[di]!r}")
