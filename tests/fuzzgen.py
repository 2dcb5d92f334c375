"""Adversarial document generator for differential tests (seeded).  Test infrastructure."""
import random

ALPH = {
    "letters": list("abcdefgstmdrvlLSTMDRVE") + ["é", "ß", "ſ", "Ж", "я", "λ", "中", "文", "ａ", "İ", "ǅ", "K"],
    "digits": list("0123456789") + ["٣", "௧", "５", "²", "½", "Ⅷ", "〇"],
    "space": [" ", " ", " ", " ", "\n", "\n", "\t", "\r", "\r\n", "\x0b", "\x0c", "\x85", "\xa0", " ", " ", " ",
              " ", " ", " ", "　"],
    "punct": list("'''.,!?-_\"()[]{}<>|/\\@#$%^&*+=~`:;") + ["’", "—", "…", "€", "😀", "🙂", "‍", "́", "​", "\x00", "\x7f",
                                                              "﻿", "\U0010ffff", "⭢", "§"],
}
CONTR = ["'s", "'t", "'re", "'ve", "'m", "'ll", "'d", "'S", "'T", "'RE", "'Ve", "'M", "'lL", "'D", "'ſ", "'r", "'l", "'v", "''s"]
WORDS = ["hello", "world", "the", "a", "I", "don", "you", "we", "naïve", "über", "привет", "мир", "日本語", "한국어", "x", "Hello",
         "token", "izer", "ing", "un", "aaa", "aaaa", "abab", "123", "42", "007"]


def rand_doc(rng, max_len=60):
    n = rng.randint(0, max_len)
    out = []
    mode = rng.random()
    for _ in range(n):
        u = rng.random()
        if mode < 0.3:  # word-ish text
            if u < 0.55: out.append(rng.choice(WORDS))
            elif u < 0.80: out.append(rng.choice(ALPH["space"][:8]))
            elif u < 0.88: out.append(rng.choice(CONTR))
            elif u < 0.94: out.append(rng.choice(ALPH["punct"]))
            else: out.append(rng.choice(ALPH["digits"]) * rng.randint(1, 7))
        else:  # char soup
            if u < 0.30: out.append(rng.choice(ALPH["letters"]))
            elif u < 0.45: out.append(rng.choice(ALPH["digits"]))
            elif u < 0.75: out.append(rng.choice(ALPH["space"]))
            elif u < 0.85: out.append("'")
            elif u < 0.97: out.append(rng.choice(ALPH["punct"]))
            else: out.append(chr(rng.choice([rng.randint(0x80, 0x2FF), rng.randint(0x300, 0x36F), rng.randint(0x370, 0xD7FF),
                                              rng.randint(0xE000, 0xFFFF), rng.randint(0x10000, 0x1FFFF)])))
    return "".join(out)


def run_doc(rng):
    """Long runs (newlines, spaces, digits, letters, punctuation) that cross chunk / window / page edges."""
    out = []
    for _ in range(rng.randint(1, 6)):
        kind = rng.random()
        n = rng.choice([1, 2, 3, 15, 16, 17, 31, 32, 33, 47, 48, 49, 63, 64, 65, 100, 130, rng.randint(1, 300)])
        if kind < 0.2: out.append(rng.choice(["\n", "\r\n", "\r", "\n\n"]) * n)
        elif kind < 0.4: out.append(rng.choice([" ", "\t", "\xa0", "\u3000", " \t"]) * n)
        elif kind < 0.55: out.append("".join(rng.choice(" \n\t\r\xa0") for _ in range(n)))
        elif kind < 0.7: out.append("".join(rng.choice("0123456789٣５") for _ in range(n)))
        elif kind < 0.8: out.append(rng.choice(["a", "é", "中", "ab"]) * n)
        elif kind < 0.9: out.append(rng.choice(["!", "'", "-", "…", "!'"]) * n)
        else: out.append(rng.choice(WORDS))
        if rng.random() < 0.5: out.append(rng.choice(["x", "!", "1", "'s", " ", "", "\n", " x", "é"]))
    return "".join(out)


def rand_docs(seed, n, max_len=60):
    rng = random.Random(seed)
    docs = [run_doc(rng) if rng.random() < 0.15 else rand_doc(rng, max_len) for _ in range(n)]
    # always include the empty doc and a few fixed nasties
    docs[: min(n, 6)] = ["", " ", "\n", "'s", "a", "  "][: min(n, 6)]
    return docs
