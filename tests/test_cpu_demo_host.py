"""CPU tier: host glue of the notebook-level boundary (styletts2_b200/text.py, demo.py) against fixtures written from the
unmodified reference by oracle/make_golden_demo.py."""
import ast
import json
import os

import torch

from util import GOLD


def test_textcleaner_table_equals_the_reference_table():
    from styletts2_b200.text import SYMBOLS, TextCleaner
    g = json.load(open(os.path.join(GOLD, "textcleaner_vocab.json"), encoding="utf-8"))
    assert SYMBOLS == g["symbols"] and len(SYMBOLS) == 178
    tc = TextCleaner()
    for text, ids in g["sample_ids"].items():
        assert tc(text) == ids
    assert tc("a0b") == tc("ab")          # unknown characters are skipped, as in the reference


def test_word_tokenize_separates_punctuation_like_the_notebooks_need():
    from styletts2_b200.text import word_tokenize
    assert word_tokenize("yes, it is.") == ["yes", ",", "it", "is", "."]
    assert " ".join(word_tokenize("a test ! ")) == "a test !"
    assert word_tokenize("") == []
    g = json.load(open(os.path.join(GOLD, "textcleaner_vocab.json"), encoding="utf-8"))
    for text in g["sample_ids"]:               # val_list rows are already tokenised: the round trip keeps them
        assert " ".join(word_tokenize(text)) == " ".join(text.split())


def test_notebook_cell_fixture_is_python_and_defines_the_entry_points():
    cells = json.load(open(os.path.join(GOLD, "notebook_cells.json"), encoding="utf-8"))
    want = {"lj_inference": "inference", "lj_LFinference": "LFinference", "libri_inference": "inference",
            "libri_LFinference": "LFinference", "libri_STinference": "STinference"}
    for k, fn in want.items():
        tree = ast.parse(cells[k]["source"])
        assert [n.name for n in tree.body if isinstance(n, ast.FunctionDef)] == [fn]


def test_demo_signatures_match_the_notebook_cells():
    """Argument names and defaults of bind(...)'s callables == those of the reference cells (parsed, not executed)."""
    import inspect

    from styletts2_b200 import demo
    from styletts2_b200.configs import MODEL_CFGS
    cells = json.load(open(os.path.join(GOLD, "notebook_cells.json"), encoding="utf-8"))

    def cell_sig(src):
        f = ast.parse(src).body[0]
        return [a.arg for a in f.args.args], [ast.literal_eval(d) for d in f.args.defaults]

    class FakeSyn:
        def __init__(self, *a, **k):
            self.device = torch.device("cpu")
    orig = demo.Synthesizer, demo.make_sampler
    demo.Synthesizer, demo.make_sampler = FakeSyn, lambda m: None
    try:
        for model, keys in (("ljspeech", {"inference": "lj_inference", "LFinference": "lj_LFinference"}),
                            ("libritts", {"inference": "libri_inference", "LFinference": "libri_LFinference",
                                          "STinference": "libri_STinference"})):
            nb = demo.bind({}, MODEL_CFGS[model], "cpu")
            for fn, cell in keys.items():
                names, defaults = cell_sig(cells[cell]["source"])
                sig = inspect.signature(nb[fn])
                assert list(sig.parameters) == names, (fn, list(sig.parameters), names)
                got = [p.default for p in sig.parameters.values() if p.default is not inspect.Parameter.empty]
                assert got == defaults, (fn, got, defaults)
    finally:
        demo.Synthesizer, demo.make_sampler = orig
