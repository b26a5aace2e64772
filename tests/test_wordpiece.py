"""mmda_b200.wordpiece (host side of SURVEY.md row N3; reference call: src/data_loader.py:80-85,
``bert_tokenizer.encode_plus``) against an independent implementation of the published BERT
tokenization: the ``tokenizers`` library's BertNormalizer + BertPreTokenizer + WordPiece pipeline
(the engine behind HF's ``BertTokenizer``), over a synthetic vocabulary -- bert-base-uncased's
``vocab.txt`` cannot be fetched offline."""
import random

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from mmda_b200.wordpiece import WordPieceTokenizer, basic_tokenize

ALPHABET = list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789") + \
    list(".,!?'\"-()$+<>^`|~@#[]{}_/\\:;") + [" ", " ", " ", "\t", "\n", "\r", " "] + \
    list("éèñüÅçôï") + list("中国語日本") + ["\x00", "\x07", "​", "�", "¿", "—", "…"]


def _vocab(seed=0):
    """specials + single characters + random multi-character pieces (whole-word and ## forms)"""
    rng = random.Random(seed)
    base = ["[PAD]", "[unused0]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    chars = list("abcdefghijklmnopqrstuvwxyz0123456789") + list(".,!?'\"-()$+<>^`|~@#[]{}_/\\:;") + \
        list("中国語") + ["¿", "—"]          # "日", "本", "…" stay out of the vocabulary -> [UNK]
    pieces = set(chars) | {"##" + c for c in "abcdefghijklmnopqrstuvwxyz0123456789"[:30]}
    letters = "abcdefghijklmnopqrstuvwxyz"
    for _ in range(400):
        w = "".join(rng.choice(letters[:8]) for _ in range(rng.randint(2, 5)))
        pieces.add(w if rng.random() < 0.5 else "##" + w)
    pieces -= {"##y", "##z"}                  # some words cannot be completed -> whole word [UNK]
    return base + sorted(pieces)


@pytest.fixture(scope="module")
def pair():
    from tokenizers import Tokenizer, models, normalizers, pre_tokenizers
    vocab = _vocab()
    mine = WordPieceTokenizer(vocab)
    ref = Tokenizer(models.WordPiece({t: i for i, t in enumerate(vocab)}, unk_token="[UNK]",
                                     max_input_chars_per_word=100))
    ref.normalizer = normalizers.BertNormalizer(clean_text=True, handle_chinese_chars=True,
                                                strip_accents=None, lowercase=True)
    ref.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    return mine, ref


def test_known_cases(pair):
    mine, _ = pair
    assert basic_tokenize("Hello, World!  don't") == ["hello", ",", "world", "!", "don", "'", "t"]
    assert basic_tokenize("Café naïve Åre") == ["cafe", "naive", "are"]
    assert basic_tokenize("ab中国cd") == ["ab", "中", "国", "cd"]
    assert basic_tokenize("a\x00b\x07c�d") == ["abcd"]
    assert basic_tokenize("") == [] and basic_tokenize(" \t\n") == []
    assert mine.tokenize("x" * 101) == ["[UNK]"]            # over max_input_chars_per_word
    assert mine.tokenize("日本") == ["[UNK]", "[UNK]"]        # CJK characters are words of their own
    assert mine.tokenize("abz") == ["[UNK]"]                 # no "##z": the WHOLE word is unknown


@settings(max_examples=600, deadline=None)
@given(st.lists(st.sampled_from(ALPHABET), min_size=0, max_size=60).map("".join))
def test_matches_the_tokenizers_library_on_fuzzed_text(pair, text):
    mine, ref = pair
    enc = ref.encode(text, add_special_tokens=False)
    assert mine.tokenize(text) == enc.tokens
    assert mine.ids(text) == enc.ids


def test_long_words_and_sentences_match(pair):
    mine, ref = pair
    rng = random.Random(5)
    for _ in range(200):
        words = ["".join(rng.choice("abcdefgh") for _ in range(rng.choice([1, 3, 8, 40, 99, 100, 101, 150])))
                 for _ in range(rng.randint(1, 30))]
        text = " ".join(words)
        assert mine.ids(text) == ref.encode(text, add_special_tokens=False).ids


def test_encode_plus_contract_of_the_reference_call(pair):
    """data_loader.py:84-85 + :108-110: [CLS] pieces [SEP], right-truncated to max_length, padded
    with [PAD]; token types all 0; mask 1 on real tokens."""
    mine, _ = pair
    text = "abc def, ghab!"
    ids = mine.ids(text)
    assert len(ids) >= 5
    for L in (2, 3, len(ids) + 1, len(ids) + 2, len(ids) + 7):
        e = mine.encode_plus(text, max_length=L, add_special_tokens=True, pad_to_max_length=True)
        body = ids[:L - 2]
        want = [mine.cls_id] + body + [mine.sep_id]
        assert e["input_ids"] == want + [mine.pad_id] * (L - len(want))
        assert e["attention_mask"] == [1] * len(want) + [0] * (L - len(want))
        assert e["token_type_ids"] == [0] * L
        assert len(e["input_ids"]) == L
    e = mine.encode_plus(text)                     # no max_length: nothing cut, nothing padded
    assert e["input_ids"] == [mine.cls_id] + ids + [mine.sep_id]
    with pytest.raises(ValueError):
        mine.encode_plus(text, max_length=1)
    with pytest.raises(KeyError):
        WordPieceTokenizer(["a", "b"])             # a vocabulary without the specials


def test_feeds_the_device_dataset_flattening(pair, tmp_path):
    """wire-format samples -> flatten_split(wordpiece_ids=tok.sample_ids): ragged id arrays the
    device collate (mmda_collate_bert) consumes; vocab file round trip."""
    from mmda_b200.collate import flatten_split
    mine, _ = pair
    p = tmp_path / "vocab.txt"
    p.write_text("\n".join(_vocab()) + "\n", encoding="utf-8")
    tok = WordPieceTokenizer.from_vocab_file(str(p))
    assert tok.vocab == mine.vocab
    rng = np.random.default_rng(0)
    texts = [["Abc", "def,"], ["gh!"], ["a", "b", "c", "ab-cd"]]
    samples = [((rng.integers(2, 50, len(w)), rng.normal(size=(len(w), 5)).astype(np.float32),
                 rng.normal(size=(len(w), 7)).astype(np.float32), w),
                np.zeros((1, 7), dtype=np.float32), f"seg{i}") for i, w in enumerate(texts)]
    flat = flatten_split(samples, wordpiece_ids=tok.sample_ids)
    want = [tok.ids(" ".join(w)) for w in texts]
    assert flat["wp_offsets"].tolist() == np.cumsum([0] + [len(x) for x in want]).tolist()
    assert flat["wp_ids"].tolist() == [i for x in want for i in x]
