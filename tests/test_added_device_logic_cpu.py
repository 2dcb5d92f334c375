"""CPU: the device code of the added-token extraction (tokenizers_b200/csrc/added_kernels.cuh: A1 candidate scan, A2
per-document resolution) compiled for the host by tests/native/added_emul.cpp and fuzzed against the host logic of
tokenizers_b200/added.py (which tests/test_added_tokens.py pins to the reference wheel): same spans, same ids, same hard /
inner / added bitmaps.  The tables are built here the way b2t_engine_set_added_tokens builds them."""
import ctypes, os, subprocess
import numpy as np
import pytest
import helpers
from oracle import oracle as orc
from tokenizers_b200 import added

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "native", "libadded_emul.so")
PAGE, MAX_SPAN = 2048, 256


def _emul():
    src = os.path.join(HERE, "native", "added_emul.cpp")
    hdr = os.path.join(helpers.ROOT, "tokenizers_b200", "csrc", "added_kernels.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        inc = "/usr/local/cuda/include"
        if not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
            pytest.skip("CUDA headers not available")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + inc, "-Wno-attributes", "-shared", "-fPIC", "-o", SO, src])
    L = ctypes.CDLL(SO)
    L.b2t_emul_added.restype = ctypes.c_uint32
    L.b2t_emul_added.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint32] + [ctypes.c_void_p] * 8 + \
        [ctypes.c_uint32, ctypes.c_void_p] + [ctypes.c_void_p] * 8 + [ctypes.c_uint32, ctypes.c_void_p]
    return L


def _tables(specs):
    """the device tables of b2t_engine_set_added_tokens (engine.cu): two sets, longest token first, first-byte / first-pair bitmaps"""
    sets = [[], []]
    for i, (content, sw, ls, rs, nm, sp) in enumerate(specs):
        sets[1 if nm else 0].append((content.encode("utf-8"), 1000 + i, (1 if sw else 0) | (2 if ls else 0) | (4 if rs else 0)))
    tb, to, ti, tf, begin = bytearray(), [0], [], [], [0, 0, 0]
    first, pair = np.zeros(16, dtype=np.uint32), np.zeros(2 * 2048, dtype=np.uint32)
    for s in range(2):
        begin[s] = len(ti)
        for b, tid, fl in sorted(sets[s], key=lambda t: -len(t[0])):   # stable: equal lengths keep their order
            tb += b; to.append(len(tb)); ti.append(tid); tf.append(fl)
            first[8 * s + (b[0] >> 5)] |= np.uint32(1 << (b[0] & 31))
            for b1 in range(256):
                if len(b) > 1 and b1 != b[1]:
                    continue
                two = b[0] | (b1 << 8)
                pair[2048 * s + (two >> 5)] |= np.uint32(1 << (two & 31))
    begin[2] = len(ti)
    fb = [b for b in range(256) if ((int(first[b >> 5]) | int(first[8 + (b >> 5)])) >> (b & 31)) & 1]
    n_first = len(fb) if len(fb) <= 4 else 0
    bc = np.zeros(4, dtype=np.uint32)
    for i, b in enumerate(fb[:4] if n_first else []):
        bc[i] = b * 0x01010101
    cls = orc.class_table("rust")
    packed = np.zeros(0x110000 // 16, dtype=np.uint32)
    for k in range(16):
        packed |= cls[k::16].astype(np.uint32) << np.uint32(2 * k)
    return dict(tb=np.frombuffer(bytes(tb) + b"\0" * 8, dtype=np.uint8).copy(), to=np.asarray(to, dtype=np.uint32), ti=np.asarray(ti, dtype=np.uint32),
                tf=np.asarray(tf + [0], dtype=np.uint8), begin=np.asarray(begin, dtype=np.uint32), first=first, pair=pair, cls=packed, n_first=n_first, bc=bc)


def _run(L, T, docs):
    data, off = helpers.pack_docs(docs)
    n, nd = int(off[-1]), len(docs)
    buf = np.zeros(n + 64, dtype=np.uint8); buf[:n] = data
    nw, npg = n // 32 + 2, n // PAGE + 1
    cand0, cand1, anyb = np.zeros(nw, np.uint32), np.zeros(nw, np.uint32), np.zeros(nw // 32 + 2, np.uint32)
    hard, inner, addedb = np.zeros(nw, np.uint32), np.zeros(nw, np.uint32), np.zeros(nw, np.uint32)
    for p in off.tolist():
        hard[int(p) >> 5] |= np.uint32(1 << (int(p) & 31))
    head = np.full(npg, 0xFFFFFFFF, dtype=np.uint32)
    cap = n // 16 + 4096
    pool = np.zeros(2 * cap, dtype=np.uint32)
    err = np.zeros(1, dtype=np.uint32)
    P = lambda a: a.ctypes.data
    used = L.b2t_emul_added(P(buf), n, P(off), nd, P(T["tb"]), P(T["to"]), P(T["ti"]), P(T["tf"]), P(T["begin"]), P(T["first"]), P(T["pair"]), P(T["cls"]),
                            T["n_first"], P(T["bc"]), P(cand0), P(cand1), P(anyb), P(hard), P(inner), P(addedb), P(head), P(pool), cap, P(err))
    bit = lambda a, p: (int(a[p >> 5]) >> (p & 31)) & 1
    spans = []
    for pg in range(npg):
        i = int(head[pg])
        while i != 0xFFFFFFFF:
            v, nxt = int(pool[2 * i]), int(pool[2 * i + 1])
            start = pg * PAGE + (v & (PAGE - 1))
            stop = start + 1
            while stop < n and bit(inner, stop):
                stop += 1
            spans.append((start, stop, v >> 11))
            i = nxt
    assert used == len(spans)
    return sorted(spans), int(err[0]), (hard, inner, addedb), off


SPEC_SETS = [helpers.ADDED_TOKEN_SPECS,
             [("<s>", False, False, False, False, True), ("</s>", False, False, True, False, True), ("<mask>", False, True, False, False, True)],
             [("ab", False, False, False, True, False), ("abc", False, False, False, True, False), ("b", True, False, False, False, False), (" x", False, False, True, False, True)]]


@pytest.mark.parametrize("si", range(len(SPEC_SETS)))
def test_device_extraction_logic_matches_host_logic(si):
    L = _emul()
    specs = SPEC_SETS[si]
    T = _tables(specs)
    entries = [{"id": 1000 + i, "content": c, "single_word": sw, "lstrip": ls, "rstrip": rs, "normalized": nm, "special": sp}
               for i, (c, sw, ls, rs, nm, sp) in enumerate(specs)]
    av = added.AddedVocabulary(entries, orc.class_table("rust"))
    checked = refused = 0
    for seed in range(6):
        docs = helpers.added_token_docs(300 + seed, 500)
        if si == 2:
            import random
            rng = random.Random(seed)
            docs = ["".join(rng.choice(["a", "b", "c", "ab", "abc", " ", "x", " x", "é", "　", "."]) for _ in range(rng.randint(0, 30))) for _ in range(800)]
        # one document at a time as well as all together: the batch form exercises spans near document and page boundaries
        for group in ([docs] + [[d] for d in docs[:60]]):
            spans, err, (hard, inner, addedb), off = _run(L, T, group)
            exp, unsupported = [], False
            for d, doc in enumerate(group):
                b = doc.encode("utf-8")
                last = 0
                for tid, a, e in av.extract(b):
                    if tid is not None:
                        if a < last or e - a > MAX_SPAN:
                            unsupported = True     # overlapping spans (the reference's rstrip corner) / a span beyond the device limit
                        exp.append((int(off[d]) + a, int(off[d]) + e, tid))
                    last = max(last, e)
            if unsupported:
                assert err & 8, group[:3]
                refused += 1
                continue
            assert err == 0 and spans == exp, (si, seed, [g for g in group][:2], spans[:5], exp[:5])
            bit = lambda a, p: (int(a[p >> 5]) >> (p & 31)) & 1
            for a, e, _ in exp[:200]:
                assert bit(addedb, a) and bit(hard, a) and (e == int(off[-1]) or bit(hard, e)) and not bit(inner, a)
            checked += len(exp)
    assert checked > 500, (checked, refused)
