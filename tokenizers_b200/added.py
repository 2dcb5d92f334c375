"""Added / special token extraction: the host step that runs BEFORE the accelerated path on every encode.

Restates `AddedVocabulary::extract_and_normalize` (tokenizer/added_vocabulary.rs:430-490, 523-564 in the reference) for
pipelines without a normalizer (the only ones the engine accepts): the text is first split on the added tokens with
`normalized == false`, then every remaining piece on the ones with `normalized == true`; matches are leftmost-longest
and non-overlapping (the reference builds an Aho-Corasick automaton with MatchKind::LeftmostLongest; here an ordered
byte-regex alternation, longest pattern first, which selects the same matches), filtered by `single_word`, widened by
`lstrip` / `rstrip`.  The pieces in between go to the GPU engine as rows of their own -- the pre-tokenizer never sees
across an added token -- and `stitch_rows` puts the rows of a document back together (token order, offsets relative
to the document, word indices counted over all splits like PreTokenizedString::into_encoding, pre_tokenizer.rs:198-263).

`\\w` / `\\s` in the single_word / lstrip / rstrip rules are the Rust `regex` crate's Unicode classes
(added_vocabulary.rs:99-102); the class table comes from libb2t.so (`b2t_unicode_class_table`, host-only call).
"""
import re
import numpy as np

CLS_W, CLS_S = 1, 3  # b2t_unicode_class_table(rust): 1 = \w, 3 = \s


class AddedToken:
    __slots__ = ("id", "content", "single_word", "lstrip", "rstrip", "normalized", "special")

    def __init__(self, d):
        self.id = int(d["id"])
        self.content = d["content"]
        self.single_word = bool(d.get("single_word", False))
        self.lstrip = bool(d.get("lstrip", False))
        self.rstrip = bool(d.get("rstrip", False))
        self.normalized = bool(d.get("normalized", not d.get("special", False)))
        self.special = bool(d.get("special", False))


def _char_before(b, end):
    """(code point, start) of the UTF-8 character that ends at byte `end` of b (end > 0)."""
    s = end - 1
    while s > 0 and (b[s] & 0xC0) == 0x80 and end - s < 4:
        s -= 1
    return _decode(b, s)[0], s


def _decode(b, s):
    """(code point, length) of the UTF-8 character starting at byte s."""
    b0 = b[s]
    if b0 < 0x80:
        return b0, 1
    n = 2 if b0 < 0xE0 else (3 if b0 < 0xF0 else 4)
    cp = b0 & (0x7F >> n)
    for k in range(1, n):
        cp = (cp << 6) | ((b[s + k] & 0x3F) if s + k < len(b) else 0)
    return cp, n


class AddedVocabulary:
    def __init__(self, entries, class_table):
        """entries: tokenizer.json's `added_tokens`; class_table: uint8[0x110000] with the Rust \\w / \\s classes."""
        self._cls = class_table
        self.tokens = {}
        by_content = {}
        for d in entries:
            if not d.get("content"):
                continue  # added_vocabulary.rs:288-291: empty tokens are ignored
            t = AddedToken(d)
            self.tokens[t.id] = t
            by_content[t.content] = t
        self.by_content = by_content
        self._sets = [self._compile([t for t in by_content.values() if not t.normalized]),
                      self._compile([t for t in by_content.values() if t.normalized])]
        # prefilter keys: the first two bytes of every token (the whole token if it has one byte); a document without any
        # of them cannot hold an added token.  bytes.find is a C loop, unlike a regex alternation over the whole buffer
        self.keys = sorted({t.content.encode("utf-8")[:2] for t in by_content.values()})
        alts = []
        for b0 in sorted({k[:1] for k in self.keys}):
            if b0 in self.keys:
                alts.append(re.escape(b0))  # a one-byte token: every occurrence of the byte is a candidate
            else:
                # the second byte is a lookahead so that every candidate START is reported, also inside runs ("   " holds
                # two candidates for the token "  ", and the first may belong to the previous document)
                alts.append(re.escape(b0) + b"(?=[" + b"".join(re.escape(k[1:2]) for k in self.keys if k[:1] == b0) + b"])")
        self.prefilter = re.compile(b"|".join(alts)) if alts else None

    @staticmethod
    def _compile(tokens):
        if not tokens:
            return None
        pats = sorted((t.content.encode("utf-8") for t in tokens), key=lambda p: (-len(p), p))
        return re.compile(b"|".join(re.escape(p) for p in pats)), {t.content.encode("utf-8"): t for t in tokens}

    def __len__(self):
        return len(self.tokens)

    def _class(self, cp):
        return int(self._cls[cp]) if cp < 0x110000 else 0

    # added_vocabulary.rs:104-125
    def _ends_with_word(self, b, end):
        return end > 0 and self._class(_char_before(b, end)[0]) == CLS_W

    def _starts_with_word(self, b, start):
        return start < len(b) and self._class(_decode(b, start)[0]) == CLS_W

    def _space_leftmost_at_end(self, b, end):
        while end > 0:
            cp, s = _char_before(b, end)
            if self._class(cp) != CLS_S:
                break
            end = s
        return end

    def _space_rightmost_at_start(self, b, start):
        p = start
        while p < len(b):
            cp, n = _decode(b, p)
            if self._class(cp) != CLS_S:
                break
            p += n
        return p - start

    def find_matches(self, sentence, which):
        """added_vocabulary.rs:430-490.  sentence: bytes.  -> [(id or None, start, stop)] covering the sentence."""
        if not sentence:
            return [(None, 0, 0)]
        ms = self._sets[which]
        if ms is None:
            return [(None, 0, len(sentence))]
        rx, by_bytes = ms
        start_offset, splits = 0, []
        for m in rx.finditer(sentence):
            start, stop = m.span()
            tok = by_bytes[m.group()]
            if tok.single_word:
                start_space = start == 0 or not self._ends_with_word(sentence, start)
                stop_space = stop == len(sentence) or not self._starts_with_word(sentence, stop)
                if not stop_space or not start_space:
                    continue
            if tok.lstrip:
                start = max(self._space_leftmost_at_end(sentence, start), start_offset)
            if tok.rstrip:
                stop += self._space_rightmost_at_start(sentence, stop)
            if start_offset < start:
                splits.append((None, start_offset, start))
            splits.append((tok.id, start, stop))
            start_offset = stop
        if start_offset != len(sentence):
            splits.append((None, start_offset, len(sentence)))
        return splits

    def extract(self, doc):
        """added_vocabulary.rs:523-564 without a normalizer.  doc: bytes -> [(id or None, start, stop)], empty pieces dropped."""
        out = []
        for tid, a, b in self.find_matches(doc, 0):
            if tid is not None:
                out.append((tid, a, b))
            elif b > a:
                for tid2, a2, b2 in self.find_matches(doc[a:b], 1):
                    if tid2 is not None or b2 > a2:
                        out.append((tid2, a + a2, a + b2))
        return out


def split_batch(av, data, doc_off):
    """Rows for the engine.  A document in which an added token occurs is cut so that each of its plain-text pieces is
    a row of its own; the bytes in between (the added tokens' spans) become filler rows, whose tokens are dropped when
    stitching, so that the rows still tile the buffer.
    -> (row_off uint64[r+1], parts, cut) with parts[d] = None for an untouched document d, else
       (first_row, n_rows, [(id or None, row index or None, start, stop)])"""
    n_docs = len(doc_off) - 1
    arr = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
    parts = [None] * n_docs
    if av is None or av.prefilter is None or arr.size == 0:
        return np.asarray(doc_off, dtype=np.uint64), parts, False
    # candidate positions: one pass of a two-byte-prefix regex over the buffer (no copy: re takes any bytes-like object;
    # its literal-prefix scan runs at memchr speed when the first bytes are rare in the text)
    pos = np.fromiter((m.start() for m in av.prefilter.finditer(memoryview(arr))), dtype=np.int64)
    if pos.size == 0:
        return np.asarray(doc_off, dtype=np.uint64), parts, False
    docs = np.unique(np.searchsorted(np.asarray(doc_off, dtype=np.int64), pos, side="right") - 1)
    cuts = {}
    for d in docs.tolist():
        a, b = int(doc_off[d]), int(doc_off[d + 1])
        sp = av.extract(arr[a:b].tobytes())
        if len(sp) == 1 and sp[0][0] is None:
            continue  # the candidate did not survive (e.g. single_word): the document stays one row
        cuts[d] = sp
    if not cuts:
        return np.asarray(doc_off, dtype=np.uint64), parts, False
    rows = [0]
    for d in range(n_docs):
        base, end = int(doc_off[d]), int(doc_off[d + 1])
        sp = cuts.get(d)
        if sp is None:
            rows.append(end)
            continue
        first = len(rows) - 1
        plist, cursor = [], 0
        # plain pieces are disjoint and ordered (find_matches); added-token spans may overlap them only in the
        # reference's rstrip corner case, which is why they are not rows themselves
        for tid, a, b in sp:
            if tid is not None:
                plist.append((tid, None, a, b))
                continue
            if a > cursor:
                rows.append(base + a)  # filler
            plist.append((None, len(rows) - 1, a, b))
            rows.append(base + b)
            cursor = b
        if base + cursor < end or len(rows) - 1 == first:
            rows.append(end)  # trailing filler (or the only row of a document that is one added token)
        parts[d] = (first, len(rows) - 1 - first, plist)
    return np.asarray(rows, dtype=np.uint64), parts, True


def stitch_rows(data, doc_off, parts, ids, offs, wid, row_ptr, byte_offsets):
    """Row CSR (engine output for `split_batch`'s rows) -> document CSR.  offs / wid may be None."""
    n_docs = len(doc_off) - 1
    arr = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
    o_ids, o_offs, o_wid = [], [], []
    added_at = []  # (index of the token in the output, absolute byte span) of every added token, for offset trimming
    rp = np.zeros(n_docs + 1, dtype=np.uint64)
    row, run_start_row, total = 0, 0, 0  # untouched documents are copied in bulk

    def flush(upto_row):
        a, b = int(row_ptr[run_start_row]), int(row_ptr[upto_row])
        if b > a:
            o_ids.append(ids[a:b])
            if offs is not None: o_offs.append(offs[a:b])
            if wid is not None: o_wid.append(wid[a:b])

    for d in range(n_docs):
        pl = parts[d]
        if pl is None:
            total += int(row_ptr[row + 1]) - int(row_ptr[row])
            row += 1
            rp[d + 1] = total
            continue
        flush(row)
        first, n_rows, plist = pl
        base = int(doc_off[d])
        words = 0
        for tid, r, a, b in plist:
            if byte_offsets:
                ua, ub = a, b
            else:  # characters = bytes that are not UTF-8 continuation bytes
                ua = int(np.count_nonzero((arr[base:base + a] & 0xC0) != 0x80))
                ub = ua + int(np.count_nonzero((arr[base + a:base + b] & 0xC0) != 0x80))
            if tid is not None:
                o_ids.append(np.array([tid], dtype=np.uint32))
                if offs is not None: o_offs.append(np.array([[ua, ub]], dtype=np.uint32))
                if wid is not None: o_wid.append(np.array([words], dtype=np.uint32))
                added_at.append((total, base + a, base + b))
                words += 1
                total += 1
                continue
            ta, tb = int(row_ptr[r]), int(row_ptr[r + 1])
            if tb > ta:
                o_ids.append(ids[ta:tb])
                if offs is not None: o_offs.append(offs[ta:tb] + np.uint32(ua))
                if wid is not None:
                    o_wid.append(wid[ta:tb] + np.uint32(words))
                    words += int(wid[tb - 1]) + 1
                total += tb - ta
        row = first + n_rows
        run_start_row = row
        rp[d + 1] = total
    flush(row)
    cat = lambda xs, shape, dt: np.concatenate(xs) if xs else np.zeros(shape, dtype=dt)
    return (cat(o_ids, 0, np.uint32), None if offs is None else cat(o_offs, (0, 2), np.uint32),
            None if wid is None else cat(o_wid, 0, np.uint32), rp, added_at)
