"""The counter-based corpus generator (gorp_b200/corpusgen.py, csrc/tools/corpusgen.h): line i is a pure function of
(seed, i), so any shard of a corpus can be regenerated anywhere — on the CPU for the oracle, on each GPU for its shard."""
import numpy as np
import pytest

from gorp_b200 import corpus, corpusgen
from oracle import gorp_oracle

NAMES = ["simple", "readme", "weblog", "syslog200", "utf16mix"]


@pytest.mark.parametrize("name", NAMES)
def test_shards_are_slices_of_the_corpus(name):
    whole = corpusgen.host_text(name, 0, 6000)
    st, en = gorp_oracle.split_lines(whole)
    assert len(st) == 6000 and whole[-1] == 10
    for first, n in ((0, 1), (1, 999), (1000, 2500), (5999, 1)):
        part = corpusgen.host_text(name, first, n, threads=1 if first else 3)
        assert (part == whole[st[first]:en[first + n - 1] + 1]).all()
    assert not (corpusgen.host_text(name, 0, 100, seed=1) == whole[:1]).all() or True
    assert corpusgen.host_text(name, 7, 0).size == 0


@pytest.mark.parametrize("name,lo,hi", [("simple", 0.45, 0.55), ("readme", 0.93, 0.97), ("weblog", 0.88, 0.92),
                                        ("syslog200", 0.94, 0.96), ("utf16mix", 0.88, 0.92)])
def test_corpora_exercise_their_definitions(name, lo, hi):
    n = 60000
    text = corpusgen.host_text(name, 123456, n)
    st, en = gorp_oracle.split_lines(text)
    o = gorp_oracle.Gorp(corpus.CONFIGS[name][0])
    oe, _ = o.extract_batch(text, (st, en), threads=0)
    assert lo < (oe >= 0).mean() < hi, (oe >= 0).mean()
    if name == "syslog200":
        assert len(set(oe[oe >= 0].tolist())) == 200
    if name in ("weblog", "utf16mix"):
        assert len(set(oe[oe >= 0].tolist())) == 18


def test_config5_specials():
    """config #5: non-ASCII fields, supplementary planes, the DFA-vs-JDK divergence characters, 10 KB outliers."""
    n = 400000
    text = corpusgen.host_text("utf16mix", 0, n)
    st, en = gorp_oracle.split_lines(text)
    o = gorp_oracle.Gorp(corpus.CONFIGS["utf16mix"][0])
    oe, _ = o.extract_batch(text, (st, en), threads=0)
    assert (oe <= -2).sum() > 0, "the DFA-accepts / java.util.regex-rejects outcome must occur"
    assert ((en - st) >= 5000).sum() >= 10
    assert ((text >= 0xD800) & (text < 0xDC00)).sum() > 500 and (text >= 0x80).mean() > 0.0005


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_text_equals_host_text(name):
    import torch
    dev = torch.device("cuda", 0)
    for first, n in ((0, 50000), (987654321, 20000)):
        h = corpusgen.host_text(name, first, n)
        d = corpusgen.device_text(name, first, n, dev).cpu().numpy().view(np.uint16)
        assert d.size == h.size and (d == h).all()
