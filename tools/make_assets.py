#!/usr/bin/env python
"""Train the benchmark tokenizers offline with the reference's own trainers (the `tokenizers` wheel).

No vocabulary files exist offline (the reference downloads gpt2-vocab.json etc. from the hub:
tokenizers/Makefile:48-56), so the three pipelines of BASELINE.json are synthesised on the synthetic corpora
of SURVEY.md §8(d) and committed gz-compressed under assets/:

  gpt2_style.json.gz   ByteLevel(add_prefix_space=False) + BPE, 50 257 entries         (configs 1, 2, 5)
  llama3_style.json.gz Sequence[Split(tiktoken regex, isolated), ByteLevel(use_regex=False)] + BPE 128 000,
                       ignore_merges=True                                              (config 3)
  wordpiece.json.gz    Whitespace + WordPiece 30 522, [UNK], "##", max_input_chars_per_word=100  (config 4)

Run in the dev container:  python tools/make_assets.py [--train-mb 64]
"""
import argparse, gzip, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import corpus
from tokenizers import Tokenizer, Regex, models, pre_tokenizers, trainers

LLAMA3_PATTERN = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*"
                  r"|\s*[\r\n]+|\s+(?!\S)|\s+")
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets")


def save(tok, name):
    s = tok.to_str()
    js = json.loads(s)
    path = os.path.join(ROOT, name + ".json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(js, ensure_ascii=False, separators=(",", ":")).encode("utf-8"))
    print(name, "vocab", tok.get_vocab_size(), "->", path, os.path.getsize(path), "bytes")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--train-mb", type=int, default=64)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    os.makedirs(ROOT, exist_ok=True)
    nbytes = a.train_mb << 20

    def docs(kind, seed):
        data, off = corpus.generate(kind, seed, 0, nbytes // 300, max_bytes=nbytes)
        return corpus.to_strings(data, off)

    if a.only in ("", "gpt2"):
        t = time.time()
        tok = Tokenizer(models.BPE())
        tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
        tr = trainers.BpeTrainer(vocab_size=50257, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False)
        tok.train_from_iterator(docs(2, 1002), tr)
        save(tok, "gpt2_style")
        print("gpt2 %.1fs" % (time.time() - t))
    if a.only in ("", "llama3"):
        t = time.time()
        tok = Tokenizer(models.BPE(ignore_merges=True))
        tok.pre_tokenizer = pre_tokenizers.Sequence([
            pre_tokenizers.Split(Regex(LLAMA3_PATTERN), "isolated"),
            pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
        tr = trainers.BpeTrainer(vocab_size=128000, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False)
        tok.train_from_iterator(docs(2, 1003), tr)
        # the trainer does not carry the model flag over
        js = json.loads(tok.to_str()); js["model"]["ignore_merges"] = True
        tok = Tokenizer.from_str(json.dumps(js))
        save(tok, "llama3_style")
        print("llama3 %.1fs" % (time.time() - t))
    if a.only in ("", "wordpiece"):
        t = time.time()
        tok = Tokenizer(models.WordPiece(unk_token="[UNK]", max_input_chars_per_word=100))
        tok.pre_tokenizer = pre_tokenizers.Whitespace()
        tr = trainers.WordPieceTrainer(vocab_size=30522, special_tokens=["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"],
                                       show_progress=False)
        tok.train_from_iterator(docs(4, 1004), tr)
        save(tok, "wordpiece")
        print("wordpiece %.1fs" % (time.time() - t))


if __name__ == "__main__":
    main()
