"""CPU: the oracle (oracle/b2t_oracle.c) against the golden vectors produced by the reference implementation and
against the known-answer vectors transcribed from the reference's own unit tests.  This is what pins the oracle."""
import json, os
import numpy as np
import pytest
import helpers
from oracle import oracle as orc


@pytest.mark.parametrize("name", helpers.GOLDEN_NAMES)
def test_oracle_matches_reference_golden(name):
    tj, cases = helpers.load_golden(name)
    o = orc.Oracle(tj)
    docs = [c["input"] for c in cases]
    got = o.encode_batch(docs)
    helpers.assert_csr_equal(got, helpers.cases_to_csr(cases), docs, f"oracle vs golden_{name}")


def _tiny(pretok, add_prefix_space):
    """A byte-alphabet BPE (no merges) behind the requested pre-tokenizer, or a WordPiece for whitespace."""
    tj, _ = helpers.load_golden("bytes_only")
    j = json.loads(tj)
    if pretok == "whitespace":
        j["pre_tokenizer"] = {"type": "Whitespace"}
        j["model"] = {"type": "WordPiece", "unk_token": "[UNK]", "continuing_subword_prefix": "##", "max_input_chars_per_word": 100,
                      "vocab": {"[UNK]": 0}}
    else:
        j["pre_tokenizer"]["add_prefix_space"] = add_prefix_space
        j["pre_tokenizer"]["use_regex"] = pretok == "bytelevel"
    return json.dumps(j)


def test_oracle_reference_kats():
    kats = json.load(open(os.path.join(helpers.GOLDEN, "reference_kats.json"), encoding="utf-8"))
    for k in kats["pretokenize"]:
        o = orc.Oracle(_tiny(k["pretok"], k["add_prefix_space"]))
        assert [list(x) for x in o.pre_tokenize(k["input"])] == k["splits"], k["source"]
    o = orc.Oracle(_tiny("bytelevel", False))
    for k in kats["byte_level_offsets"]:
        ids, offs, wid, rp = o.encode_batch([k["input"]])
        assert offs.tolist() == k["offsets"] and wid.tolist() == k["word_ids"], k["source"]


def test_oracle_byte_offsets_mode():
    # OffsetType::Byte (Rust encode_batch): spans of whole original chars in bytes
    o = orc.Oracle(_tiny("bytelevel", False))
    ids, offs, wid, rp = o.encode_batch(["i⭢j é"], offset_type=orc.OFF_BYTE)
    assert offs.tolist() == [[0, 1], [1, 4], [1, 4], [1, 4], [4, 5], [5, 6], [6, 8], [6, 8]]
