#!/usr/bin/env python
"""Generates tests/golden/golden_added_tokens.json.gz from the reference wheel (tokenizers 0.22.2): added-token
extraction + special-token template on top of the asset tokenizers.  Run in the dev container: python tests/golden/make_golden_added.py"""
import gzip, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import asset_json, with_added_tokens, added_token_docs  # noqa: E402
import tokenizers  # noqa: E402


def flat(encs):
    return [{"ids": list(e.ids), "offsets": [list(o) for o in e.offsets], "word_ids": list(e.word_ids),
             "type_ids": list(e.type_ids), "special": list(e.special_tokens_mask), "tokens": list(e.tokens)} for e in encs]


configs = []
for asset, prefix_space, template, seed in [("gpt2_style", False, False, 21), ("gpt2_style", True, True, 22), ("llama3_style", None, True, 23),
                                            ("wordpiece", None, True, 24)]:
    js = json.loads(asset_json(asset))
    if prefix_space is not None:
        js["pre_tokenizer"]["add_prefix_space"] = prefix_space
    tj = with_added_tokens(json.dumps(js), template)
    tok = tokenizers.Tokenizer.from_str(tj)
    docs = added_token_docs(seed, 300)
    configs.append({"asset": asset, "prefix_space": prefix_space, "template": template, "docs": docs,
                    "expected": {str(sp): flat(tok.encode_batch(docs, add_special_tokens=sp)) for sp in (False, True)}})
out = os.path.join(HERE, "golden_added_tokens.json.gz")
with gzip.open(out, "wb") as f:
    f.write(json.dumps({"generator": "tests/golden/make_golden_added.py", "tokenizers": tokenizers.__version__, "configs": configs}).encode("utf-8"))
print(out, os.path.getsize(out))
