"""Pins the oracle (oracle/*.py, oracle/gorp_oracle.c) against every golden vector the reference's
own tests hold for the hot path. CPU only."""
import numpy as np
import pytest

from oracle import brics, frontend, gorp_oracle, jdkre
from tests import reference_vectors as V


def test_multipattern_accept_lists():
    m = brics.PolyMatcher(V.MULTI_PATTERNS)
    for s, want in V.MULTI_CASES:
        assert m.match(jdkre.to_units(s)) == want, s


@pytest.mark.parametrize("case", V.POLY_DSL)
def test_polymatch_from_dsl(case):
    g = gorp_oracle.Gorp(case[0])
    for s, want in case[1]:
        assert g.matcher.match(jdkre.to_units(s)) == want, s


@pytest.mark.parametrize("case", V.FULL_EXACT)
def test_full_extraction_exact(case):
    g = gorp_oracle.Gorp(case[0])
    for s, want in case[1]:
        assert g.extract_map(s, "id") == want, s


@pytest.mark.parametrize("case", V.FULL_SUBSET)
def test_full_extraction_subset(case):
    g = gorp_oracle.Gorp(case[0])
    for s, want in case[1]:
        got = g.extract_map(s, "id")
        assert got is not None, s
        for k, v in want.items():
            assert got[k] == v, (s, k)


def test_translator_kats():
    for a, b in V.QUOTE_KATS:
        assert frontend.quote_literal_as_regexp(a) == b
    for a, b in V.AUTOM_KATS:
        assert frontend.massage_regexp_for_automaton(a) == b
    for a, b in V.JDK_KATS:
        assert frontend.massage_regexp_for_jdk(a) == b


@pytest.mark.parametrize("case", V.DERIVED_STRINGS)
def test_generated_strings(case):
    xs = frontend.definition_to_strings(case[0])
    assert len(xs) == len(case[1])
    for x, (name, autom, jdk, names) in zip(xs, case[1]):
        assert (x.name, x.autom, x.jdk, x.extractor_names) == (name, autom, jdk, names)


@pytest.mark.parametrize("case", V.ERROR_KATS)
def test_definition_errors(case):
    with pytest.raises(frontend.DefinitionParseError) as ei:
        gorp_oracle.Gorp(case[0])
    msg = str(ei.value).lower()
    for sub in case[1]:
        assert sub.lower() in msg, (sub, msg)


def test_pattern_resolution_kat():  # PatternResolutionTest.java:11-38
    r = frontend.DefinitionReader("pattern %a a\npattern %b b\npattern %c stuff!\n"
                                  "pattern %abba (%a%b %'b'-%a)\npattern %full %abba %c\n")
    r.read_uncooked()
    r.resolve_patterns()
    got = {k: v.text for k, v in r.cooked_patterns.items()}
    assert got == {"a": "a", "b": "b", "c": "stuff!", "abba": "(ab b-a)", "full": "(ab b-a) stuff!"}


def test_template_resolution_kat():  # TemplateResolutionTest.java:44-92
    r = frontend.DefinitionReader("pattern %a a\ntemplate @base (%a:foo)\n"
                                  "template @full @base...%{[.*{2}]}--%a\n")
    r.read_uncooked()
    r.resolve_patterns()
    r.resolve_templates()
    parts = [(type(p).__name__, p.text) for p in r.cooked_templates["full"].parts]
    assert parts == [("LiteralText", "("), ("LiteralPattern", "a"), ("LiteralText", ":foo)"),
                     ("LiteralText", "..."), ("LiteralPattern", "[.*{2}]"), ("LiteralText", "--"),
                     ("LiteralPattern", "a")]


def test_uncooked_tokenising_kats():  # UncookedDefTest.java:62-211
    r = frontend.DefinitionReader("pattern %wsChar \\s\npattern %optws %wsChar*%%\npattern %word ([a-z]+)\n"
                                  "pattern %phrase3   %word %word2%word3\n")
    r.read_uncooked()
    p = r.patterns["optws"].parts
    assert [(type(x).__name__, getattr(x, "name", getattr(x, "text", None))) for x in p] == \
        [("PatternRef", "wsChar"), ("LiteralPattern", "*%")]
    p = r.patterns["phrase3"].parts
    assert [(type(x).__name__, getattr(x, "name", getattr(x, "text", None))) for x in p] == \
        [("PatternRef", "word"), ("LiteralPattern", " "), ("PatternRef", "word2"), ("PatternRef", "word3")]
    r = frontend.DefinitionReader("template @actual value=$value(Accepted$$%{\\d+})\n")
    r.read_uncooked()
    parts = r.templates["actual"].parts
    assert parts[0].text == "value=" and parts[1].name == "value"
    assert [(type(x).__name__, x.text) for x in parts[1].parts] == \
        [("LiteralText", "Accepted$"), ("LiteralPattern", "\\d+")]


def test_line_reader_kat():  # io/InputLineReaderTest.java:13-36
    stuff = ["line 1", "line 2", "# commentary", "   ", "line 3\\", " with continuation \\", "or two...\\",
             " or three!", "    # more comments"]
    sb = ""
    for i, s in enumerate(stuff, 1):
        sb += s + {0: "\r\n", 2: "\r"}.get(i % 3, "\n")
    lr = frontend._LineReader(sb)
    out = []
    while True:
        ln = lr.next_line()
        if ln is None:
            break
        out.append(ln[1])
    assert out == ["line 1", "line 2", "line 3 with continuation or two... or three!"]


def test_append_kat():  # ExtractionResolutionTest.java:14-40
    g = gorp_oracle.Gorp("pattern %a a\ntemplate @base (%a:foo)\nextract rule1 {  \n"
                         "  template @base value=$MyValue(%a:%{\\w+})\n  append { \"enabled\" : true, \"x\" : 3 }\n}")
    assert g.extractions[0].append == {"enabled": True, "x": 3}
    assert g.extractions[0].autom_source == "\\(a:foo\\)[ \t]+value=(a:[a-zA-Z_0-9]+)"
    assert g.extractions[0].regexp_source == "\\(a:foo\\)[ \t]+value=(a:\\w+)"


def test_config_table_sizes():
    """README definition -> 31 product states, 15 classes (throwaway BFS of the survey)."""
    g = gorp_oracle.Gorp(V.README_DEF)
    a = g.matcher.automata
    # 31 states; `_stride` = 34 brics start points, which collapse to the survey's 15 distinct columns
    T = a.transitions.reshape(a.n_states, a.stride)
    assert (a.n_states, a.stride, len({tuple(T[:, c]) for c in range(a.stride)})) == (31, 34, 15)
    assert [x.regexp_source for x in g.extractions][0] == "\\[(\\d+)\\]:[ \t]+(PUT)[ \t]+(\\d+)ms[ \t]+(\\S+)"
    line = "[102456879]: GET 123ms /rest-service/v1/endpoint?foo=bar"
    assert g.matcher.match(jdkre.to_units(line)) == [1, 2]
    assert g.extract_map(line, "id") == {"id": "GetRequest", "timestamp": "102456879", "verb": "GET",
                                         "timeTakenInMsec": "123", "path": "/rest-service/v1/endpoint?foo=bar",
                                         "marker": "EXTRACTED"}
    assert g.extract("102456879: GET 123ms 200 /rest-service/v1/endpoint?foo=bar")[0] == -1  # README.md:109
    g1 = gorp_oracle.Gorp(V.SIMPLE_GRP)
    assert g1.extractions[0].regexp_source == "\\<\\d+\\>(\\S+)[ \t]+\\((Accepted)\\)[ \t]+"
    assert g1.extract_map("<86>2015-05-12T20:57:53+00:00 (Accepted) ") == \
        {"eventTimeStamp": "2015-05-12T20:57:53+00:00", "authStatus": "Accepted"}
    assert g1.extract("<86>2015-05-12T20:57:53+00:00 (Accepted)")[0] == -1


def test_c_hot_loop_agrees_with_python():
    """oracle/gorp_oracle.c vs the readable Python restatement, incl. divergence inputs."""
    for d, cases in [V.FULL_SIMPLE, V.FULL_INTERMEDIATE, V.FULL_FULL, V.PARAM_EXTRACTOR, V.PARAM_TEMPLATE,
                     (V.README_DEF, []), (V.SIMPLE_GRP, [])]:
        g = gorp_oracle.Gorp(d)
        lines = [c[0] for c in cases] + [
            "", "x", "[1]: GET 2ms /a", "[1]: GET 2ms /a\x0b", "[1]: GET 2ms /a\x08b", "[1]:\x0bGET 2ms /a",
            "[12]: PUT 5ms /\U0001F600/x", "[12]: POST 5ms /\ud83d", "[12]: HEAD 77ms /p\u0085q",
            "<1>ts (Accepted) ", "<1>t\x0bs (Accepted) ", "<1>\U00010000 (Accepted) \t ", "value=foobar",
        ]
        units = [np.asarray(jdkre.to_units(s), dtype=np.uint16) for s in lines]
        text = np.concatenate([np.concatenate((u, [10])) for u in units]).astype(np.uint16)
        starts, ends = gorp_oracle.split_lines(text)
        assert len(starts) == len(lines)
        ext, spans = g.extract_batch(text, (starts, ends), threads=2)
        for i, s in enumerate(lines):
            e, sp = g.extract(s)
            assert ext[i] == e, (s, ext[i], e)
            if e >= 0:
                flat = [x for ab in sp for x in ab]
                assert spans[i, :len(flat)].tolist() == flat, s
