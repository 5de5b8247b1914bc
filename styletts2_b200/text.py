"""Host-side text glue of the Demo notebooks: symbol table -> token ids, and a word tokenizer stand-in.

`TextCleaner` maps every character of an IPA phoneme string to its index in the reference's 178-entry symbol table
(text_utils.py:3-26; the table is the tokenizer vocabulary the checkpoints were trained with, i.e. an interface
constant like the state-dict keys).  The table below is checked entry by entry against the reference's
(tests/golden/textcleaner_vocab.json, written by oracle/make_golden_demo.py from the imported reference).
"""
from __future__ import annotations

import re
from typing import Dict, List

PAD = "$"
PUNCTUATION = ';:,.!?¡¿—…"«»“” '
LETTERS = "ABCDEFGHIJKLMNOPQRSTUVWXYZ" + "abcdefghijklmnopqrstuvwxyz"
IPA = ("ɑɐɒæɓʙβɔɕçɗɖðʤəɘɚɛɜɝɞɟʄɡɠɢʛɦɧħɥʜɨɪʝɭɬɫɮʟɱɯɰŋɳɲɴøɵɸθœɶʘɹɺɾɻʀʁɽʂʃʈʧʉʊʋⱱʌɣɤʍχʎʏʑʐʒʔʡʕʢǀǁǂǃˈˌːˑʼʴʰʱʲʷˠˤ˞"
       "↓↑→↗↘'̩'ᵻ")
SYMBOLS: List[str] = [PAD, *PUNCTUATION, *LETTERS, *IPA]          # 178 entries (the apostrophe appears twice)
SYMBOL_TO_ID: Dict[str, int] = {}
for _i, _s in enumerate(SYMBOLS):
    SYMBOL_TO_ID[_s] = _i                                           # a repeated symbol keeps its LAST index, as in the reference


class TextCleaner:
    """text_utils.py:15-26: characters -> ids; unknown characters are skipped (the reference prints the text)."""

    def __init__(self, dummy=None, verbose=False):
        self.word_index_dictionary = SYMBOL_TO_ID
        self.verbose = verbose

    def __call__(self, text: str) -> List[int]:
        out = []
        for ch in text:
            i = self.word_index_dictionary.get(ch)
            if i is None:
                if self.verbose:
                    print(text)
                continue
            out.append(i)
        return out


_TOKEN = re.compile(r"\.\.\.|[^\W_]+(?:['ˈˌːʼ̩][^\W_]+)*|[^\w\s]", re.UNICODE)


def word_tokenize(text: str) -> List[str]:
    """Stand-in for nltk.word_tokenize as the notebooks use it (`' '.join(word_tokenize(ps))`: punctuation split from
    the neighbouring word).  nltk is not installed in this image; phoneme strings only need this separation rule.
    Pass the real nltk function to demo.bind(word_tokenize=...) where it is available."""
    return _TOKEN.findall(text)
