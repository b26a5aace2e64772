"""Word-piece tokenizer for the BERT text branch (host side of SURVEY.md row N3).

The reference tokenizes inside ``collate_fn`` with ``bert_tokenizer.encode_plus(text,
max_length=SENT_LEN+2, add_special_tokens=True, pad_to_max_length=True)``
(src/data_loader.py:16, :80-85) where ``bert_tokenizer`` is
``BertTokenizer.from_pretrained('bert-base-uncased')`` -- a third-party algorithm (transformers,
unpinned; ``encode_plus`` no longer exists in the 5.x series installed here, so the reference's
own collate does not run against it).  This file restates the published BERT tokenization
(google-research/bert ``tokenization.py``: BasicTokenizer + WordpieceTokenizer, uncased) from a
``vocab.txt`` and offers the two call shapes the path needs:

* ``encode_plus(...)`` with the reference's argument names and result keys, so the object can
  stand in for ``bert_tokenizer`` in ``data_loader.py``;
* ``sample_ids(sample)`` -- the word-piece ids of one wire-format sample, the callback
  ``mmda_b200.collate.DeviceDataset(..., wordpiece_ids=...)`` takes (specials, truncation and
  padding then happen on the device in ``mmda_collate_bert``).

String processing stays on the host by design: it runs once per split, not once per step.
Checked against the ``tokenizers`` library's BERT pipeline (the engine behind HF's tokenizer) on
fuzzed text in tests/test_wordpiece.py.
"""
from __future__ import annotations

import unicodedata
from typing import Dict, Iterable, List, Optional, Sequence

MAX_CHARS_PER_WORD = 100        # longer "words" become [UNK] (tokenization.py: max_input_chars_per_word)


def _is_whitespace(ch: str) -> bool:
    return ch in " \t\n\r" or unicodedata.category(ch) == "Zs"


def _is_control(ch: str) -> bool:
    if ch in "\t\n\r":
        return False
    return unicodedata.category(ch).startswith("C")


def _is_punctuation(ch: str) -> bool:
    cp = ord(ch)
    # all non-letter / non-digit ASCII counts as punctuation ("^", "$", "`" included)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def _is_cjk(cp: int) -> bool:
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or
            0x2A700 <= cp <= 0x2B73F or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF or
            0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F)


def basic_tokenize(text: str, lower: bool = True) -> List[str]:
    """Clean -> isolate CJK characters -> whitespace split -> lower-case + strip accents ->
    split every punctuation character off."""
    cleaned = []
    for ch in text:
        cp = ord(ch)
        if cp == 0 or cp == 0xFFFD or _is_control(ch):
            continue
        if _is_whitespace(ch):
            cleaned.append(" ")
        elif _is_cjk(cp):
            cleaned.append(" " + ch + " ")
        else:
            cleaned.append(ch)
    out: List[str] = []
    for tok in "".join(cleaned).split():
        if lower:
            tok = tok.lower()
            tok = "".join(c for c in unicodedata.normalize("NFD", tok)
                          if unicodedata.category(c) != "Mn")
        word: List[str] = []
        for ch in tok:
            if _is_punctuation(ch):
                if word:
                    out.append("".join(word))
                    word = []
                out.append(ch)
            else:
                word.append(ch)
        if word:
            out.append("".join(word))
    return out


class WordPieceTokenizer:
    """Uncased BERT tokenizer over a ``vocab.txt`` (one token per line, line number = id)."""

    def __init__(self, vocab: Sequence[str] | Dict[str, int], lower: bool = True,
                 unk: str = "[UNK]", cls: str = "[CLS]", sep: str = "[SEP]", pad: str = "[PAD]"):
        self.vocab: Dict[str, int] = dict(vocab) if isinstance(vocab, dict) else \
            {t: i for i, t in enumerate(vocab)}
        for special in (unk, cls, sep, pad):
            if special not in self.vocab:
                raise KeyError(f"vocabulary has no {special} entry")
        self.lower = lower
        self.unk, self.cls, self.sep, self.pad = unk, cls, sep, pad
        self.unk_id, self.cls_id = self.vocab[unk], self.vocab[cls]
        self.sep_id, self.pad_id = self.vocab[sep], self.vocab[pad]

    @classmethod
    def from_vocab_file(cls, path: str, **kw) -> "WordPieceTokenizer":
        with open(path, encoding="utf-8") as f:
            return cls([line.rstrip("\n") for line in f], **kw)

    # -- tokenization.py::WordpieceTokenizer.tokenize: greedy longest-match-first ------------
    def _pieces(self, word: str) -> List[str]:
        if len(word) > MAX_CHARS_PER_WORD:
            return [self.unk]
        pieces, start = [], 0
        while start < len(word):
            end, cur = len(word), None
            while start < end:
                sub = word[start:end] if start == 0 else "##" + word[start:end]
                if sub in self.vocab:
                    cur = sub
                    break
                end -= 1
            if cur is None:
                return [self.unk]          # one unmatched stretch makes the whole word unknown
            pieces.append(cur)
            start = end
        return pieces

    def tokenize(self, text: str) -> List[str]:
        out: List[str] = []
        for word in basic_tokenize(text, self.lower):
            out.extend(self._pieces(word))
        return out

    def ids(self, text: str) -> List[int]:
        v = self.vocab
        return [v[t] for t in self.tokenize(text)]

    # -- the reference's call (src/data_loader.py:84-85) -------------------------------------
    def encode_plus(self, text: str, max_length: Optional[int] = None, add_special_tokens: bool = True,
                    pad_to_max_length: bool = False) -> Dict[str, List[int]]:
        """``[CLS] pieces [SEP]`` truncated from the right to ``max_length`` and padded with
        ``[PAD]``; keys ``input_ids`` / ``token_type_ids`` / ``attention_mask`` as the reference
        reads them (data_loader.py:108-110)."""
        ids = self.ids(text)
        extra = 2 if add_special_tokens else 0
        if max_length is not None:
            if max_length < extra:
                raise ValueError(f"max_length {max_length} leaves no room for the special tokens")
            ids = ids[:max_length - extra]
        if add_special_tokens:
            ids = [self.cls_id] + ids + [self.sep_id]
        mask = [1] * len(ids)
        if pad_to_max_length and max_length is not None:
            fill = max_length - len(ids)
            ids, mask = ids + [self.pad_id] * fill, mask + [0] * fill
        return {"input_ids": ids, "token_type_ids": [0] * len(ids), "attention_mask": mask}

    # -- the device-resident pipeline's callback ---------------------------------------------
    def sample_ids(self, sample) -> List[int]:
        """Word-piece ids (no specials) of one wire-format sample ``((words, visual, acoustic,
        actual_words), label, segment)``: the text is ``" ".join(actual_words)`` exactly as
        data_loader.py:83 builds it."""
        return self.ids(" ".join(sample[0][3]))

    def batch_ids(self, texts: Iterable[str]) -> List[List[int]]:
        return [self.ids(t) for t in texts]
