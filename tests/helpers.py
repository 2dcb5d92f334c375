"""Shared test helpers (golden loading, CSR flattening, wheel access)."""
import gzip, json, os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ASSETS = os.path.join(ROOT, "assets")

GOLDEN_NAMES = ["gpt2", "gpt2_prefix", "llama3", "wordpiece", "bytes_only", "nonmonotone", "ignore_merges"]


def asset_json(name):
    return gzip.open(os.path.join(ASSETS, name + ".json.gz")).read().decode("utf-8")


def _set_path(d, dotted, value):
    ks = dotted.split(".")
    for k in ks[:-1]:
        d = d[k]
    d[ks[-1]] = value


def load_golden(name):
    """-> (tokenizer_json_str, cases)"""
    g = json.loads(gzip.open(os.path.join(GOLDEN, f"golden_{name}.json.gz")).read().decode("utf-8"))
    t = g["tokenizer"]
    if isinstance(t, str) and t.startswith("asset:"):
        tj = asset_json(t[6:])
    elif isinstance(t, dict) and "asset" in t:
        j = json.loads(asset_json(t["asset"]))
        for k, v in t["patch"].items():
            _set_path(j, k, v)
        tj = json.dumps(j)
    else:
        tj = json.dumps(t)
    return tj, g["cases"]


def cases_to_csr(cases):
    ids = np.array([i for c in cases for i in c["ids"]], dtype=np.uint32)
    offs = np.array([o for c in cases for o in c["offsets"]], dtype=np.uint32).reshape(-1, 2)
    wid = np.array([w for c in cases for w in c["word_ids"]], dtype=np.uint32)
    rp = np.zeros(len(cases) + 1, dtype=np.uint64)
    if cases:
        np.cumsum([len(c["ids"]) for c in cases], out=rp[1:])
    return ids, offs, wid, rp


def pack_docs(docs):
    bs = [d.encode("utf-8") for d in docs]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        np.cumsum([len(b) for b in bs], out=off[1:])
    return np.frombuffer(b"".join(bs), dtype=np.uint8).copy(), off


def assert_csr_equal(got, exp, docs=None, what=""):
    names = ["ids", "offsets", "word_ids", "row_ptr"]
    for g, e, nm in zip(got, exp, names):
        if g is None:
            continue
        if not np.array_equal(np.asarray(g).reshape(-1), np.asarray(e).reshape(-1)):
            msg = f"{what}: {nm} differ"
            if docs is not None:
                grp, erp = np.asarray(got[3]), np.asarray(exp[3])
                for d in range(len(docs)):
                    a, b = int(erp[d]), int(erp[d + 1])
                    ga, gb = (int(grp[d]), int(grp[d + 1])) if d + 1 < len(grp) else (0, 0)
                    same = (gb - ga == b - a) and all(
                        x is None or np.array_equal(np.asarray(x)[ga:gb], np.asarray(y)[a:b]) for x, y in zip(got[:3], exp[:3]))
                    if not same:
                        msg += f"\n first differing doc {d}: {docs[d]!r}\n  exp ids {np.asarray(exp[0])[a:b].tolist()} off {np.asarray(exp[1])[a:b].tolist()}" \
                               f"\n  got ids {np.asarray(got[0])[ga:gb].tolist()} off {None if got[1] is None else np.asarray(got[1])[ga:gb].tolist()}"
                        break
            raise AssertionError(msg)


def wheel():
    """The reference implementation, if importable here (it is in the dev container and on the GPU image)."""
    try:
        import tokenizers
        return tokenizers
    except Exception:
        return None


def wheel_csr(tok, docs):
    encs = tok.encode_batch(docs, add_special_tokens=False)
    return cases_to_csr([{"ids": e.ids, "offsets": e.offsets, "word_ids": e.word_ids} for e in encs])
