"""Third-party fuzz partners for the oracle AND the product's host compiler (SURVEY.md §8c "secondary cross-checks").

Both engines gorp delegates to (brics automaton 1.11-8, java.util.regex) are absent here, so the oracle restates them.
To keep the oracle from being pinned only by its author's reading, two independent implementations of the same
regex families are used as partners on randomly generated patterns of the common subset:

  * `interegular` (FSM library): per-regex LANGUAGE equivalence of the DFA dialect — the oracle's brics restatement
    (oracle/brics.py) and the product's C++ builder (gorp_compile_patterns tables) must accept exactly the strings the
    interegular FSM accepts, including `{n}`, `{n,}`, `{n,m}`, `|`, nested groups, classes, negated classes, `.`.
  * Python `re` (a leftmost, greedy/lazy backtracking engine of the same family as java.util.regex): group SPANS of
    `fullmatch` on the JDK-dialect string the front-end generates — the oracle's backtracker (oracle/jdkre.py) and
    the product's determinised Pike VM (host tables interpreted by tests/host/hosttest.cpp, incl. the tail automata)
    must give the same spans, including lazy quantifiers, alternation order, counted repeats and nested captures.

Not authoritative for Java corner cases: the generators stay inside the subset where the three
dialects agree by documentation (no `\\b`, anchors, possessive quantifiers, empty loops, surrogates, line terminators).
"""
import itertools
import re

import numpy as np
import pytest

from gorp_b200 import Blob, DefinitionReader, UnsupportedDefinition
from oracle import brics, gorp_oracle, jdkre
from tests import hostlib
from tests.test_host_compile import pack

interegular = pytest.importorskip("interegular")

ALPHA = "abc:"
STRINGS = ["".join(t) for n in range(0, 7) for t in itertools.product(ALPHA, repeat=n)]


def random_regex(rng, depth=0, top=True):
    """A regex of the subset common to brics (flags NONE), java.util.regex, Python re and interegular."""
    def atom():
        r = rng.random()
        if r < 0.40 or depth >= 3:
            return str(rng.choice(list("abc:")))
        if r < 0.50:
            return "."
        if r < 0.62:
            items = sorted(set(rng.choice(list("abc:"), size=rng.integers(1, 4)).tolist()))
            return "[" + ("^" if rng.random() < 0.3 else "") + "".join(items) + "]"
        if r < 0.68:
            return "[a-c]" if rng.random() < 0.5 else "[^b-c]"
        return "(" + random_regex(rng, depth + 1, top=False) + ")"

    def piece():
        a = atom()
        r = rng.random()
        if a[0] == "(" and any(ch in a for ch in "*+{"):
            # no unbounded repeat of a group that repeats inside: the backtracking partners (Python re) go exponential
            return a + ("?" if r > 0.8 else "")
        if r < 0.55:
            return a
        if r < 0.65:
            return a + "?"
        if r < 0.75:
            return a + "*"
        if r < 0.85:
            return a + "+"
        if r < 0.90:
            return a + "{%d}" % rng.integers(1, 4)
        if r < 0.95:
            return a + "{%d,}" % rng.integers(0, 3)
        lo = int(rng.integers(0, 3))
        return a + "{%d,%d}" % (lo, lo + int(rng.integers(0, 3)))

    def seq():
        return "".join(piece() for _ in range(rng.integers(1, 5 if depth == 0 else 3)))

    alts = [seq() for _ in range(1 if rng.random() < 0.6 else rng.integers(2, 4))]
    return "|".join(alts)


def accepts_tables(cm, tr, accept_lists, s):
    p = 0
    for c in s:
        p = tr[p, cm[ord(c)]]
        if p < 0:
            return []
    return accept_lists[p]


def test_dfa_language_vs_interegular_and_re():
    """Random sets of DFA-dialect regexes: oracle brics DFA == product tables == interegular FSM == re.fullmatch on every
    string over a 4-letter alphabet up to length 6 and on random longer strings."""
    rng = np.random.default_rng(20261017)
    longer = ["".join(rng.choice(list(ALPHA), size=rng.integers(7, 16))) for _ in range(400)]
    n_sets = 40
    for _ in range(n_sets):
        pats = [random_regex(rng) for _ in range(int(rng.integers(1, 6)))]
        fsms = [interegular.parse_pattern(p).to_fsm() for p in pats]
        res = [re.compile(p) for p in pats]
        dfas = [brics.to_min_dfa(p) for p in pats]
        blob = Blob.from_patterns(pats)
        cm, tr, af, al = blob.tables()
        for s in STRINGS + longer:
            want = [i for i, f in enumerate(fsms) if f.accepts(s)]
            assert want == [i for i, r in enumerate(res) if r.fullmatch(s)], (pats, s)
            got_oracle = []
            for i, d in enumerate(dfas):
                st = 0
                for c in s:
                    st = d.step(st, ord(c))
                    if st < 0:
                        break
                if st >= 0 and d.accept[st]:
                    got_oracle.append(i)
            assert got_oracle == want, ("oracle brics restatement", pats, s, got_oracle, want)
            got_product = list(accepts_tables(cm, tr, al, s))
            assert got_product == want, ("product tables", pats, s, got_product, want)


def random_capture_definition(rng):
    """One extraction whose template is a sequence of extractors over inline patterns (optionally nested, optionally
    separated by literal ':'), i.e. capturing groups only where gorp itself generates them (Gorp.java:94-129)."""
    def pat():
        for _ in range(20):
            p = random_regex(rng, depth=1, top=False)
            if "{" in p and "}" not in p:
                continue
            return p
        return "a"

    def lazy(p):  # sprinkle lazy quantifiers (java.util.regex and Python re agree on them)
        out = []
        for i, ch in enumerate(p):
            out.append(ch)
            if ch in "*+?" and (i == 0 or p[i - 1] not in "*+?") and rng.random() < 0.25 and (i + 1 == len(p) or p[i + 1] not in "*+?{"):
                out.append("?")
        return "".join(out)

    names = iter("xyzuvw")
    parts = []
    for _ in range(int(rng.integers(1, 4))):
        if rng.random() < 0.2:
            parts.append("$%s($%s(%%{%s})$%s(%%{%s}))" % (next(names), next(names), lazy(pat()), next(names), lazy(pat())))
            break
        parts.append("$%s(%%{%s})" % (next(names), lazy(pat())))
        if rng.random() < 0.3:
            parts.append(":")
    return "extract e {\n template " + "".join(parts) + "\n}\n"


def test_capture_spans_vs_python_re():
    """Random extraction definitions: for every short string (and random longer ones) the oracle's Gorp.extract outcome and
    spans equal what Python's re.fullmatch gives on the SAME generated JDK-dialect string, and the product's compiled
    tables (general path and tail automata) equal the oracle."""
    rng = np.random.default_rng(424242)
    longer = ["".join(rng.choice(list(ALPHA), size=rng.integers(7, 14))) for _ in range(300)]
    lines = STRINGS[:1365] + longer  # every string up to length 5
    done = refused = 0
    while done < 60 and refused < 400:
        definition = random_capture_definition(rng)
        try:
            o = gorp_oracle.Gorp(definition)
            g = DefinitionReader.reader(definition).read()
        except (UnsupportedDefinition, jdkre.Unsupported):
            refused += 1  # nullable loops, stacked quantifiers, ...: outside the supported subset on purpose
            continue
        x = o.extractions[0]
        try:
            jdk = re.compile(x.regexp_source)
            # brics has no lazy quantifiers: postfix operators stack, so `x+?` reads as `(x+)?` = `x*`, `x*?` as `(x*)?` = `x*`,
            # `x??` as `(x?)?` = `x?` — the DFA dialect accepts MORE than java.util.regex for `+?` (such lines end as
            # CAPTURE_FAIL in the reference). The partner gets the same language spelled without stacked operators.
            fsm = interegular.parse_pattern(x.autom_source.replace("+?", "*").replace("*?", "*").replace("??", "?")).to_fsm()
        except Exception:  # noqa: BLE001  (a dialect-specific escape the partners do not read the same way)
            refused += 1
            continue
        G = len(x.extractor_names)
        text, starts, ends = pack(lines)
        oe, osp = o.extract_batch(text, (starts, ends), threads=1)
        for i, s in enumerate(lines):
            m = jdk.fullmatch(s)
            dfa_ok = fsm.accepts(s)
            want = -1 if not dfa_ok else (0 if m else -2)
            assert oe[i] == want, (definition, s, int(oe[i]), want)
            if want == 0:
                spans = [v for k in range(1, G + 1) for v in m.span(k)]
                assert osp[i][:2 * G].tolist() == spans, (definition, s, osp[i][:2 * G].tolist(), spans)
        ext, spans, _ = hostlib.run(g.blob().bytes(), text, starts, ends, 2 * G)
        assert (ext == oe).all() and (spans == osp[:, :2 * G]).all(), definition
        r = hostlib.run_tails(g.blob().bytes(), text, 2 * G)
        if r is not None:
            assert (r[0] == oe).all() and (r[1] == osp[:, :2 * G]).all(), definition
        done += 1
    assert done >= 40, (done, refused)
