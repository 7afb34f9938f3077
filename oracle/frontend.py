"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of gorp's definition front-end: the code that turns a `.grp`
definition into, per extraction, (name, extractor names, automaton-dialect
regex string, JDK-dialect regex string, append JSON).

Follows (all under /root/reference/gorp-core/src/main/java/com/salesforce/gorp/):
  io/InputLineReader.java:69-150      physical -> logical lines
  util/TokenHelper.java:25-276        keyword / name / inline-pattern tokenising
  DefinitionReader.java:126-640       declarations, template tokenising
  model/CookedDefinitions.java:57-475 pattern / template / extraction resolution
  util/RegexHelper.java:20-237        the three dialect translators
  Gorp.java:50-129                    pieces -> the two regex strings

Reference quirks that are reproduced on purpose (they are observable):
  * a template referenced BEFORE its declaration resolves to an EMPTY template
    (CookedDefinitions.java:235-237 iterates the new, empty CookedTemplate);
  * a duplicate extraction name replaces the body but keeps the first position
    (UncookedDefinitions.java:42-45, LinkedHashMap);
  * whitespace runs in literal text that START with space/TAB become `[ \\t]+`
    with a literal TAB inside the class (RegexHelper.java:28-34).
"""
from __future__ import annotations

import json
import re


class DefinitionParseError(Exception):
    """Mirrors DefinitionParseException / IOException of the reference."""


# --------------------------------------------------------------------------
# io/InputLineReader.java
# --------------------------------------------------------------------------

def _physical_lines(text: str):
    """BufferedReader.readLine: terminators are \\n, \\r, \\r\\n only
    (InputLineReader.java:110)."""
    out, i, n, start = [], 0, len(text), 0
    while i < n:
        c = text[i]
        if c == "\n" or c == "\r":
            out.append(text[start:i])
            if c == "\r" and i + 1 < n and text[i + 1] == "\n":
                i += 1
            i += 1
            start = i
        else:
            i += 1
    if start < n:
        out.append(text[start:])
    return out


def _is_empty_or_comment(line: str) -> bool:
    # InputLineReader.java:140-150
    for ch in line:
        if ord(ch) <= 0x20:
            continue
        return ch == "#"
    return True


class _LineReader:
    def __init__(self, text: str, src: str = "<input string>"):
        self.lines = _physical_lines(text)
        self.pos = 0
        self.row = 0
        self.src = src

    def error(self, msg):
        raise DefinitionParseError("(%s, row %d): %s" % (self.src, self.row, msg))

    def next_line(self):
        """InputLineReader.java:69-96. Returns (row, contents) or None."""
        line = None
        while self.pos < len(self.lines):
            cand = self.lines[self.pos]
            self.pos += 1
            self.row += 1
            if not _is_empty_or_comment(cand):
                line = cand
                break
        if line is None:
            return None
        start = self.row
        if not line.endswith("\\"):
            return (start, line)
        combo = line[:-1]
        while True:
            # continuation lines are NOT comment/blank filtered (:84-85)
            if self.pos >= len(self.lines):
                self.error("Unexpected end-of-input when expecting line continuation'")
            seg = self.lines[self.pos]
            self.pos += 1
            self.row += 1
            if not seg.endswith("\\"):
                return (start, combo + seg)
            combo += seg[:-1]


# --------------------------------------------------------------------------
# util/TokenHelper.java
# --------------------------------------------------------------------------

def _is_ws(c):  # TokenHelper.java:259-261
    return c <= " "


def _is_num(c):
    return "0" <= c <= "9"


def _is_ident_start(c):
    # Character.isJavaIdentifierStart: letters, '_', '$' (currency), ...
    return c == "$" or c == "_" or c.isalpha()


def _is_ident_part(c):
    return _is_ident_start(c) or c.isdigit()


_KEYWORD = re.compile(r"[ \t\n\x0B\f\r]*([a-zA-Z_0-9]*)[ \t\n\x0B\f\r]*(.*)", re.S)


def _find_keyword(contents):
    # TokenHelper.java:17,25-33 — \s*(\w*)\s*(.*) with matches(); '.' excludes
    # line terminators but a logical line holds none.
    m = _KEYWORD.fullmatch(contents)
    if not m:
        return None
    return m.group(1), m.start(2)


def _find_type_marker(marker, contents, ix):
    # TokenHelper.java:41-53
    while ix < len(contents):
        c = contents[ix]
        if c == marker:
            return ix
        if _is_ws(c):
            break
        ix += 1
    return -1


def _skip_space(contents, ix):
    while ix < len(contents) and _is_ws(contents[ix]):
        ix += 1
    return ix


def _match_remaining(contents, ix, ch):
    # TokenHelper.java:82-98
    found = False
    end = len(contents)
    while ix < end:
        c = contents[ix]
        ix += 1
        if c == ch:
            if found:
                break
            found = True
        elif not _is_ws(c):
            break
    return ix if found else -1


def _parse_if_nonneg_number(s):
    if not s:
        return -1
    num = 0
    for c in s:
        if not _is_num(c):
            return -1
        num = num * 10 + (ord(c) - 48)
    return num


class _Ctx:
    """Error reporting context for one logical line (InputLine.reportError)."""

    def __init__(self, row, contents, src="<input string>"):
        self.row, self.contents, self.src = row, contents, src

    def error(self, col, fmt, *args):
        msg = fmt % args if args else fmt
        raise DefinitionParseError("[%s (%d,%d)]: %s" % (self.src, self.row, col, msg))


def _parse_name(kind, ctx, contents, ix, allow_numbers):
    # TokenHelper.java:140-191
    end = len(contents)
    if ix >= end:
        ctx.error(end, "Missing %s name", kind)
    c = contents[ix]
    name = None
    if c == '"' or c == "'":
        ix += 1
        q = contents.find(c, ix)
        if q < 0:
            ctx.error(end, "Missing closing quote ('%s') for %s name", c, kind)
        name = contents[ix:q]
        rest = q + 1
    elif not _is_ident_start(c):
        if _is_num(c):
            if allow_numbers:
                st = ix
                while ix < end and _is_num(contents[ix]):
                    ix += 1
                name = contents[st:ix]
            else:
                ctx.error(ix, "Invalid variable reference instead of %s name: can not use variable "
                              "references here (missing parenthesis after template name?)", kind)
        rest = ix
    else:
        st = ix
        ix += 1
        while ix < end and _is_ident_part(contents[ix]):
            ix += 1
        name = contents[st:ix]
        rest = ix
    return name, rest


def _parse_name_skip_space(kind, ctx, contents, ix):
    # TokenHelper.java:112-133
    name, rest = _parse_name(kind, ctx, contents, ix, False)
    end = len(contents)
    if rest >= end:
        return name, rest
    if not _is_ws(contents[rest]):
        ctx.error(rest, "Missing space character after %s name '%s'", kind, name)
    rest += 1
    while rest < end and _is_ws(contents[rest]):
        rest += 1
    return name, rest


def _parse_inline_pattern(ctx, contents, start):
    # TokenHelper.java:197-221
    end = len(contents)
    nesting, i = 1, start
    while i < end:
        c = contents[i]
        i += 1
        if c == "\\":
            i += 1
            continue
        if c == "{":
            nesting += 1
        elif c == "}":
            nesting -= 1
            if nesting == 0:
                return contents[start:i - 1], i
    ctx.error(start, "Missing closing '{' for inline pattern")


# --------------------------------------------------------------------------
# model/* pieces
# --------------------------------------------------------------------------

class Piece:
    __slots__ = ("ctx", "off")


class LiteralText(Piece):
    def __init__(self, ctx, off, text):
        self.ctx, self.off, self.text = ctx, off, text


class LiteralPattern(Piece):
    def __init__(self, ctx, off, text):
        self.ctx, self.off, self.text = ctx, off, text


class PatternRef(Piece):
    def __init__(self, ctx, off, name):
        self.ctx, self.off, self.name = ctx, off, name


class TemplateRef(Piece):
    def __init__(self, ctx, off, name, params=None):
        self.ctx, self.off, self.name, self.params = ctx, off, name, params

    def append(self, p):  # TemplateReference.java:57-62
        if self.params is None:
            self.params = []
        self.params.append(p)


class TemplateParamRef(Piece):
    def __init__(self, ctx, off, parent, pos):
        self.ctx, self.off, self.parent, self.pos = ctx, off, parent, pos


class ExtractorParamRef(Piece):
    def __init__(self, ctx, off, parent, pos):
        self.ctx, self.off, self.parent, self.pos = ctx, off, parent, pos


class Extractor(Piece):
    def __init__(self, ctx, off, name, pos=-1, parts=None):
        self.ctx, self.off, self.name, self.pos = ctx, off, name, pos
        self.parts = [] if parts is None else parts

    def append(self, p):
        self.parts.append(p)

    def with_name(self, name):  # ExtractorExpression.java: keeps parts, clears position
        return Extractor(self.ctx, self.off, name, -1, self.parts)

    def with_parts(self, parts):
        return Extractor(self.ctx, self.off, self.name, self.pos, parts)


class _ParamCollector:  # model/ParameterCollector.java
    def __init__(self):
        self.types = []

    def add(self, ctx, off, pos, typ):
        pos -= 1
        while len(self.types) <= pos:
            self.types.append("\0")
        old = self.types[pos]
        if old != typ and old != "\0":
            ctx.error(off, "Inconsistent references to parameter %d: %s vs %s", pos + 1, old, typ)
        self.types[pos] = typ


class _Uncooked:  # model/UncookedDefinition.java
    def __init__(self, ctx, name, has_params, def_start):
        self.ctx, self.name, self.def_start = ctx, name, def_start
        self.params = _ParamCollector() if has_params else None
        self.parts = []

    def append(self, p):
        self.parts.append(p)


class _CookedTemplate:
    def __init__(self, unc):
        self.name = unc.name
        self.parts = []
        self.param_types = None if unc.params is None else "".join(unc.params.types)

    def append(self, p):
        self.parts.append(p)


class FlattenedExtraction:
    def __init__(self, name, parts, extractor_names, append):
        self.name, self.parts, self.extractor_names, self.append = name, parts, extractor_names, append


# --------------------------------------------------------------------------
# DefinitionReader.java
# --------------------------------------------------------------------------

class DefinitionReader:
    KNOWN = "(pattern, template, extract)"
    PROPS = "(template, append)"

    def __init__(self, text, src="<input string>"):
        self.lr = _LineReader(text, src)
        self.src = src
        self.patterns = {}     # name -> _Uncooked   (insertion ordered)
        self.templates = {}
        self.extractions = {}  # name -> (template _Uncooked, append dict|None, append_json)
        self._read = False
        self.cooked_patterns = {}
        self.cooked_templates = {}
        self.flattened = None

    # ---- readUncooked (:126-181)
    def read_uncooked(self):
        if self._read:
            return
        self._read = True
        while True:
            ln = self.lr.next_line()
            if ln is None:
                break
            ctx = _Ctx(ln[0], ln[1], self.src)
            contents = ln[1]
            kw = _find_keyword(contents)
            if kw is None:
                ctx.error(0, "No keyword found from line; expected one of %s", self.KNOWN)
            word, rest = kw
            if word == "pattern":
                self._read_pattern(ctx, rest)
            elif word == "template":
                self._read_template(ctx, rest)
            elif word == "extract":
                self._read_extraction(ctx, rest)
            else:
                ctx.error(0, 'Unrecognized keyword "%s" encountered; expected one of %s', word, self.KNOWN)
        for p in self.patterns.values():
            self._tokenize_pattern(p)
        for t in self.templates.values():
            self._tokenize_template(t.ctx, t.def_start, t, -1,
                                    "template '%s' definition" % t.name, t.params)
        for (t, _a, _j) in self.extractions.values():
            self._tokenize_template(t.ctx, t.def_start, t, 0,
                                    "extraction template for '%s'" % t.name, None)

    def _read_pattern(self, ctx, offset):  # :189-207
        contents = ctx.contents
        ix = _find_type_marker("%", contents, offset)
        if ix < 0:
            ctx.error(offset, "Pattern name must be prefixed with '%'")
        offset = ix + 1
        name, rest = _parse_name_skip_space("pattern", ctx, contents, offset)
        if name in self.patterns:
            ctx.error(offset, "Duplicate pattern definition for name '%s'", name)
        self.patterns[name] = _Uncooked(ctx, name, False, rest)

    def _tokenize_pattern(self, unp):  # :209-258
        ctx, contents = unp.ctx, unp.ctx.contents
        end = len(contents)
        offset = unp.def_start
        ix = contents.find("%", offset)
        if ix < 0:
            unp.append(LiteralPattern(ctx, offset, contents[offset:]))
            return
        sb = []
        if ix > 0:
            sb.append(contents[offset:ix])
        while ix < end:
            c = contents[ix]
            ix += 1
            if c != "%":
                sb.append(c)
                continue
            if ix == end:
                ctx.error(ix, "Orphan '%%' at end of pattern '%s' definition", unp.name)
            c = contents[ix]
            if c == "%":
                sb.append(c)
                ix += 1
                continue
            ref, rest = _parse_name("pattern", ctx, contents, ix, False)
            if sb and "".join(sb):
                unp.append(LiteralPattern(ctx, offset, "".join(sb)))
            sb = []
            unp.append(PatternRef(ctx, ix, ref))
            ix = rest
        if sb and "".join(sb):
            unp.append(LiteralPattern(ctx, offset, "".join(sb)))

    def _read_template(self, ctx, start):  # :260-294
        contents = ctx.contents
        ix = _find_type_marker("@", contents, start)
        if ix < 0:
            ctx.error(start, "Template name must be prefixed with '@'")
        ix += 1
        name, rest = _parse_name("template", ctx, contents, ix, False)
        name_off = ix
        ix = rest
        has_params = False
        if ix + 1 < len(contents) and contents[ix] == "(" and contents[ix + 1] == ")":
            ix += 2
            has_params = True
        ix2 = _skip_space(contents, ix)
        if ix == ix2:
            ctx.error(ix, "Missing space character after template name '%s'", name)
        ix = ix2
        if name in self.templates:
            ctx.error(name_off, "Duplicate template definition for name '%s'", name)
        self.templates[name] = _Uncooked(ctx, name, has_params, ix)

    def _tokenize_template(self, ctx, ix, container, paren_count, desc, vars_):  # :303-392
        contents = ctx.contents
        end = len(contents)
        sb = []
        lit_start = ix
        got_vars = vars_ is not None
        while ix < end:
            c = contents[ix]
            ix += 1
            if c in "%@$":
                if ix == end:
                    ctx.error(ix, "Orphan '%s' at end of %s", c, desc)
                d = contents[ix]
                if c == d:
                    sb.append(c)
                    ix += 1
                    continue
                if sb:
                    container.append(LiteralText(ctx, lit_start, "".join(sb)))
                    sb = []
                if c == "%":
                    if d == "{":
                        ix += 1
                        pat, rest = _parse_inline_pattern(ctx, contents, ix)
                        container.append(LiteralPattern(ctx, ix, pat))
                    else:
                        nm, rest = _parse_name("pattern", ctx, contents, ix, False)
                        container.append(PatternRef(ctx, ix, nm))
                    ix = rest
                elif c == "@":
                    ix = self._tokenize_template_ref(ctx, ix, desc, vars_, container)
                else:
                    nm, rest = _parse_name("extractor", ctx, contents, ix, got_vars)
                    ix = rest
                    pos = _parse_if_nonneg_number(nm) if (got_vars and nm is not None) else -1
                    if got_vars and pos >= 0:
                        if pos < 1 or pos > 999999:
                            ctx.error(ix, "Invalid extractor name parameter %d in %s", pos, desc)
                        vars_.add(ctx, ix, pos, "$")
                        extr = Extractor(ctx, ix, str(pos), pos)
                    else:
                        extr = Extractor(ctx, ix, nm)
                    container.append(extr)
                    # _tokenizeInlineExtractor (:510-526)
                    if ix >= end or contents[ix] != "(":
                        ctx.error(ix, "Invalid declaration for extractor '%s': missing opening parenthesis",
                                  extr.name)
                    ix += 1
                    ix = self._tokenize_template(ctx, ix, extr, 1,
                                                 "extractor '%s' expression" % extr.name, vars_)
                lit_start = ix
                continue
            if paren_count > 0:
                if c == "(":
                    paren_count += 1
                elif c == ")":
                    paren_count -= 1
                    if paren_count == 0:
                        break
            sb.append(c)
        if sb:
            container.append(LiteralText(ctx, lit_start, "".join(sb)))
        if paren_count > 0:
            ctx.error(ix, "Missing closing parenthesis at end of %s", desc)
        return ix

    def _tokenize_template_ref(self, ctx, ix, desc, vars_, container):  # :445-479
        contents = ctx.contents
        ident, rest = _parse_name("template parameter", ctx, contents, ix, vars_ is not None)
        ix = rest
        pos = _parse_if_nonneg_number(ident) if (vars_ is not None and ident is not None) else -1
        if vars_ is not None and pos >= 0:
            if pos < 1 or pos > 999999:
                ctx.error(ix, "Invalid template parameter %d in %s", pos, desc)
            vars_.add(ctx, ix, pos, "@")
            container.append(TemplateParamRef(ctx, ix, getattr(container, "name", None), pos))
        else:
            refd = self.templates.get(ident)
            if refd is None:
                ctx.error(ix, "Referencing non-existing template '@%s' from '%s'", ident, desc)
            ref = TemplateRef(ctx, ix, ident)
            container.append(ref)
            if refd.params is not None:
                ix = self._tokenize_param_list(ctx, ix, desc, vars_, ref)
        return ix

    def _tokenize_param_list(self, ctx, ix, desc, vars_, ref):  # :394-443
        contents = ctx.contents
        end = len(contents)
        if ix >= end or contents[ix] != "(":
            ctx.error(ix, "Missing parameter list for template reference '@%s'", ref.name)
        ix += 1
        idx = 1
        while ix < end:
            c = contents[ix]
            ix += 1
            if c == ")":
                return ix
            if idx > 1:
                if c != ",":
                    ctx.error(ix, "Unexpected character %r in template parameter list for '@%s': "
                                  "expected either ',' or ')')'", c, ref.name)
                if ix >= end:
                    break
                c = contents[ix]
                ix += 1
            if c == "@":
                ix = self._tokenize_template_ref(ctx, ix, desc, vars_, ref)
            elif c == "$":
                # _tokenizeExtractorParameter (:481-504)
                ident, rest = _parse_name("extractor parameter", ctx, contents, ix, vars_ is not None)
                ix = rest
                pos = _parse_if_nonneg_number(ident) if (vars_ is not None and ident is not None) else -1
                if vars_ is not None and pos >= 0:
                    if pos < 1 or pos > 999999:
                        ctx.error(ix, "Invalid extractor parameter %d in %s", pos, desc)
                    vars_.add(ctx, ix, pos, "$")
                    ref.append(ExtractorParamRef(ctx, ix, ref.name, pos))
                else:
                    ref.append(Extractor(ctx, ix, ident))
            else:
                ctx.error(ix, "Unexpected character %r in template parameter list for '@%s': expected "
                              "either type marker '@' or closing ')'", c, ref.name)
            idx += 1
        ctx.error(ix, "Unexpected end of line within parameter list for template '@%s'", ref.name)

    def _read_extraction(self, ctx, offset):  # :528-594
        contents = ctx.contents
        name, rest = _parse_name_skip_space("extraction", ctx, contents, offset)
        ix = _match_remaining(contents, rest, "{")
        if ix != len(contents):
            ctx.error(rest, "Unexpected content for extraction '%s': expected only opening '{'", name)
        template = None
        append = None
        append_raw = []
        while True:
            ln = self.lr.next_line()
            if ln is None:
                self.lr.error("Unexpected end-of-input in extraction '%s' definition" % name)
            ctx = _Ctx(ln[0], ln[1], self.src)
            contents = ln[1]
            ix = _match_remaining(contents, 0, "}")
            if ix >= 0:
                if ix >= len(contents):
                    break
                ctx.error(rest, "Unexpected content after closing '}' for extraction '%s'", name)
            ix = _skip_space(contents, 0)
            prop, ix = _parse_name_skip_space("extraction", ctx, contents, ix)
            if prop == "template":
                if template is not None:
                    ctx.error(ix, "More than one 'template' specified for '%s'" % name)
                template = _Uncooked(ctx, "", False, ix)
            elif prop == "append":
                raw = contents[ix:].strip()
                if raw:
                    if not raw.startswith("{") and raw.startswith('"'):
                        raw = "{" + raw + "}"
                    try:
                        val = json.loads(raw)
                    except Exception as e:  # noqa: BLE001
                        ctx.error(ix, "Invalid JSON content to 'append': %s", e)
                    if not isinstance(val, dict):
                        ctx.error(ix, "Invalid 'append' value: must be JSON Object, or sequence of "
                                      "key/value pairs; was parsed as %s", type(val).__name__)
                    append_raw.append(raw)
                    if append is None:
                        append = val
                    else:
                        append.update(val)
            else:
                ctx.error(ix, 'Unrecognized extraction property "%s" encountered; expected one of %s',
                          prop, self.PROPS)
        if template is None:
            ctx.error(ix, "Missing 'template' for extraction '%s'", name)
        # LinkedHashMap.put: replacing keeps the original insertion position
        self.extractions[name] = (template, append, append_raw)

    # ---- CookedDefinitions.resolvePatterns (:57-132)
    def resolve_patterns(self):
        for name, p in self.patterns.items():
            if name in self.cooked_patterns:
                continue
            self.cooked_patterns[name] = self._resolve_pattern(name, p, None)

    def _resolve_pattern(self, name, d, stack):
        pieces = d.parts
        if len(pieces) == 1:
            p = pieces[0]
            if isinstance(p, LiteralPattern):
                return p
            if stack is None:
                stack = []
            return self._resolve_pattern_ref(name, p, stack)
        sb = []
        for p in pieces:
            if isinstance(p, LiteralPattern):
                lit = p
            else:
                if stack is None:
                    stack = []
                lit = self._resolve_pattern_ref(name, p, stack)
            sb.append(lit.text)
        off = pieces[0].off if pieces else 0
        return LiteralPattern(d.ctx, off, "".join(sb))

    @staticmethod
    def _stack_desc(marker, stack, last):
        if stack is None:
            return ""
        return "(" + "".join(marker + s + "->" for s in stack) + marker + last + ")"

    def _resolve_pattern_ref(self, from_name, ref, stack):
        to = ref.name
        res = self.cooked_patterns.get(to)
        if res is not None:
            return res
        stack.append(from_name)
        if to in stack:
            ref.ctx.error(ref.off, "Cyclic pattern reference to '%%%s' %s", to,
                          self._stack_desc("%", stack, to))
        raw = self.patterns.get(to)
        if raw is None:
            ref.ctx.error(ref.off, "Referencing non-existing pattern '%%%s' %s", to,
                          self._stack_desc("%", stack, to))
        p = self._resolve_pattern(to, raw, stack)
        self.cooked_patterns[to] = p
        stack.pop()
        return p

    # ---- resolveTemplates (:144-242)
    def resolve_templates(self):
        for name, t in self.templates.items():
            if name in self.cooked_templates:
                continue
            result = _CookedTemplate(t)
            self._resolve_template_contents(self.templates, t.name, t.parts, result, None, name)
            self.cooked_templates[name] = result

    def _resolve_template_contents(self, unc_templates, name, to_resolve, result, stack, top):
        for d in list(to_resolve):
            if isinstance(d, (LiteralText, LiteralPattern)):
                result.append(d)
            elif isinstance(d, PatternRef):
                p = self.cooked_patterns.get(d.name)
                if p is None:
                    d.ctx.error(d.off, "Referencing non-existing pattern '%%%s' from template '%s' %s",
                                d.name, top, self._stack_desc("@", stack, getattr(result, "name", "")))
                result.append(p)
            elif isinstance(d, TemplateRef):
                if d.params is not None:
                    result.append(d)
                else:
                    if stack is None:
                        stack = []
                    tmpl = self._resolve_template_ref(unc_templates, name, d, stack, top)
                    for p in tmpl.parts:
                        result.append(p)
            elif isinstance(d, Extractor):
                resolved = d.with_parts([])
                if stack is None:
                    stack = []
                self._resolve_template_contents(unc_templates, name, d.parts, resolved, stack, top)
                result.append(resolved)
            elif isinstance(d, TemplateParamRef):
                result.append(d)
            else:
                d.ctx.error(0, "Internal error: unexpected definition type %s when resolving template "
                               "definition '%s'", type(d).__name__, top)

    def _resolve_template_ref(self, unc_templates, from_name, ref, stack, top):
        to = ref.name
        res = self.cooked_templates.get(to)
        if res is not None:
            return res
        stack.append(from_name)
        if to in stack:
            ref.ctx.error(ref.off, "Cyclic template reference to '%%%s' %s", to,
                          self._stack_desc("@", stack, to))
        raw = unc_templates.get(to)
        if raw is None:
            ref.ctx.error(ref.off, "Referencing non-existing template '%%%s' %s", to,
                          self._stack_desc("@", stack, to))
        result = _CookedTemplate(raw)
        # Reference quirk (CookedDefinitions.java:235-237): the contents that get
        # resolved are result.getParts() — the NEW template's own, empty list —
        # so a forward-referenced template is cached EMPTY.
        self._resolve_template_contents(unc_templates, result.name, result.parts, result, stack, top)
        self.cooked_templates[to] = result
        stack.pop()
        return result

    # ---- resolveExtractions (:255-453)
    def resolve_extractions(self):
        self.flattened = []
        for xname, (raw_t, append, append_raw) in self.extractions.items():
            template = _CookedTemplate(raw_t)
            self._resolve_template_contents({}, raw_t.name, raw_t.parts, template, None, raw_t.name)
            names = []
            parts = []
            self._resolve_extraction_parts(template.parts, parts, names, None, top=True)
            self.flattened.append(FlattenedExtraction(xname, parts, names, append))
            self.flattened[-1].append_raw = append_raw

    def _resolve_literal(self, part, parts):
        if isinstance(part, (LiteralText, LiteralPattern)):
            parts.append(part)
            return True
        if isinstance(part, PatternRef):
            p = self.cooked_patterns.get(part.name)
            if p is None:
                raise DefinitionParseError("Internal error: non-existing pattern '%%%s', should have "
                                           "been caught earlier" % part.name)
            parts.append(p)
            return True
        return False

    def _resolve_extractor(self, part, parts, names, bindings):
        if not isinstance(part, Extractor):
            return False
        extr = part
        if extr.pos >= 0:
            p = bindings[extr.pos - 1] if (bindings is not None and 1 <= extr.pos <= len(bindings)) else None
            if not isinstance(p, Extractor):
                part.ctx.error(part.off, "Internal error: unexpected extractor parameter of type %s "
                                         "(expecting ExtractorExpression)", type(p).__name__)
            if p.pos >= 0:
                part.ctx.error(part.off, "Internal error: positional extractor parameter (%d) resolves "
                                         "to another positional (%d)", extr.pos, p.pos)
            extr = extr.with_name(p.name)
        if extr.name in names:
            part.ctx.error(part.off, "Duplicate extractor name ($%s)", extr.name)
        names.append(extr.name)
        new_parts = []
        self._resolve_extraction_parts(extr.parts, new_parts, names, bindings)
        parts.append(extr.with_parts(new_parts))
        return True

    def _resolve_extraction_parts(self, input_parts, result, names, bindings, top=False):
        for part in input_parts:
            if isinstance(part, TemplateParamRef):
                if top:
                    part.ctx.error(part.off, "Internal error: should not encounter template parameter "
                                             "%s#%d", part.parent, part.pos)
                if bindings is None:
                    part.ctx.error(part.off, "Invalid parameter variable reference @%d; template takes "
                                             "no parameters", part.pos)
                param = bindings[part.pos - 1] if 1 <= part.pos <= len(bindings) else None
                if param is None:
                    part.ctx.error(part.off, "Invalid parameter variable reference @%d; template takes "
                                             "%d parameters", part.pos, len(bindings))
                part = param
            if self._resolve_literal(part, result) or self._resolve_extractor(part, result, names, bindings):
                continue
            if isinstance(part, TemplateRef):
                self._resolve_template_ref_from_extraction(part, result, names, bindings)
                continue
            part.ctx.error(part.off, "Internal error: unrecognized DefPiece %s", type(part).__name__)

    def _resolve_template_ref_from_extraction(self, ref, result, names, incoming):  # :301-341
        template = self.cooked_templates.get(ref.name)
        if template is None:
            ref.ctx.error(ref.off, "Internal error: reference to unknown template '@%s'", ref.name)
        bindings = None
        if template.param_types is not None:
            params = ref.params or []
            pcount = len(template.param_types)
            if len(params) != pcount:
                ref.ctx.error(ref.off, "Parameter mismatch: template '@%s' expects %d parameters; %d passed",
                              ref.name, pcount, len(params))
            bindings = []
            for i, piece in enumerate(params):
                exp = template.param_types[i]
                ok = isinstance(piece, TemplateRef) if exp == "@" else (
                    isinstance(piece, Extractor) if exp == "$" else None)
                if ok is None:
                    raise DefinitionParseError("Internal error: unrecognized template parameter type %r" % exp)
                if not ok:
                    ref.ctx.error(ref.off, "Parameter mismatch: template '@%s' expects type '%s' parameter, "
                                           "got %s", ref.name, exp, type(piece).__name__)
                bindings.append(self._resolve_parameters(piece, incoming))
        self._resolve_extraction_parts(template.parts, result, names, bindings)

    def _resolve_parameters(self, piece, bindings):  # :343-381
        if isinstance(piece, TemplateParamRef):
            v = bindings[piece.pos - 1] if (bindings is not None and 1 <= piece.pos <= len(bindings)) else None
            if v is None:
                piece.ctx.error(piece.off, "Invalid parameter variable reference @%d; template has %d parameters",
                                piece.pos, 0 if bindings is None else len(bindings))
            return v
        if isinstance(piece, TemplateRef):
            if piece.params is None:
                return piece
            return TemplateRef(piece.ctx, piece.off, piece.name,
                               [self._resolve_parameters(p, bindings) for p in piece.params])
        if isinstance(piece, Extractor):
            return piece.with_parts([self._resolve_parameters(p, bindings) for p in piece.parts])
        piece.ctx.error(piece.off, "Internal error: unexpected template parameter type %s", type(piece).__name__)

    # ---- read (:74-84)
    def read(self):
        self.read_uncooked()
        if not self.extractions:
            raise DefinitionParseError("No extraction definitions found from definition")
        self.resolve_patterns()
        self.resolve_templates()
        self.resolve_extractions()
        return self.flattened


# --------------------------------------------------------------------------
# util/RegexHelper.java
# --------------------------------------------------------------------------

CC_d = "0-9"
CC_s = " \b\f\n\r\t"
CC_w = "a-zA-Z_0-9"
_QUOTE = set("()[]\\{}|*?+$^<>\"&")


def quote_literal_as_regexp(text: str) -> str:  # RegexHelper.java:20-70
    sb = []
    i, end = 0, len(text)
    while i < end:
        c = text[i]
        i += 1
        if c == " " or c == "\t":
            while i < end and text[i] <= " ":
                i += 1
            sb.append("[ \t]+")
        elif c == ".":
            sb.append("\\.")
        elif c in _QUOTE:
            sb.append("\\" + c)
        else:
            sb.append(c)
    return "".join(sb)


def _is_alpha_or_digit(d: str) -> bool:
    # Character.isAlphabetic(d) || Character.isDigit(d)
    return d.isalpha() or d.isdigit()


def massage_regexp_for_automaton(pattern: str) -> str:  # RegexHelper.java:79-201
    if "\\" not in pattern:
        return pattern
    sb = []
    end = len(pattern)
    levels = 0
    i = 0
    while i < end:
        c = pattern[i]
        i += 1
        if c == "[":
            sb.append(c)
            levels += 1
            continue
        if c == "]":
            sb.append(c)
            levels -= 1
            continue
        if c != "\\" or i >= end:
            sb.append(c)
            continue
        had_bracket = levels > 0 and pattern[i - 2] == "["
        d = pattern[i]
        i += 1
        cls = None
        if d == "\\":
            pass
        elif d == "b":
            d = "\b"
        elif d == "f":
            d = "\f"
        elif d == "n":
            d = "\n"
        elif d == "r":
            d = "\r"
        elif d == "t":
            d = "\t"
        elif d == "d":
            cls = CC_d
        elif d == "D":
            cls = "^" + CC_d
        elif d == "s":
            cls = CC_s
        elif d == "S":
            cls = "^" + CC_s
        elif d == "w":
            cls = CC_w
        elif d == "W":
            cls = "^" + CC_w
        elif _is_alpha_or_digit(d):
            raise ValueError("Unrecognized backslash escape '\\%s; can only escape backslash (\\\\), use known "
                             "control-codes (\\n, \\r, \\t), escape non-alphanumeric (\\$, \\(, ...) or refer to "
                             "a 'well-known' character class (\\s, \\S, \\d, \\D, \\w, \\W)" % d)
        if cls is not None:
            if levels == 0:
                sb.append("[" + cls + "]")
            else:
                if cls.startswith("^") and not had_bracket:
                    raise ValueError("Can not use negated character class \\%s within character class in "
                                     "position other than first (Automaton limitation)" % d)
                sb.append(cls)
            continue
        sb.append(c)
        sb.append(d)
    return "".join(sb)


def massage_regexp_for_jdk(pattern: str) -> str:  # RegexHelper.java:210-237
    sb = []
    i, end = 0, len(pattern)
    while i < end:
        c = pattern[i]
        i += 1
        if c == "\\":
            sb.append(c)
            if i < end:
                sb.append(pattern[i])
                i += 1
        elif c == "(":
            sb.append("(?:")
        else:
            sb.append(c)
    return "".join(sb)


# --------------------------------------------------------------------------
# Gorp.java:50-129 — pieces -> the two strings
# --------------------------------------------------------------------------

def _build(part, autom, jdk):
    if isinstance(part, LiteralPattern):
        try:
            autom.append(massage_regexp_for_automaton(part.text))
            jdk.append(massage_regexp_for_jdk(part.text))
        except ValueError as e:
            part.ctx.error(part.off, "Invalid pattern definition, problem (%s): %s",
                           "java.lang.IllegalArgumentException", e)
        return
    if isinstance(part, LiteralText):
        q = quote_literal_as_regexp(part.text)
        autom.append(q)
        jdk.append(q)
        return
    if isinstance(part, Extractor):
        autom.append("(")
        jdk.append("(")
        for p in part.parts:
            _build(p, autom, jdk)
        autom.append(")")
        jdk.append(")")
        return
    part.ctx.error(part.off, "Unrecognized DefPiece in FlattenedExtraction: %s", type(part).__name__)


class ExtractionStrings:
    def __init__(self, name, extractor_names, autom, jdk, append, append_raw):
        self.name, self.extractor_names = name, extractor_names
        self.autom, self.jdk, self.append, self.append_raw = autom, jdk, append, append_raw


def definition_to_strings(text: str):
    """DefinitionReader.read() up to the two regex strings per extraction."""
    flat = DefinitionReader(text).read()
    out = []
    for fx in flat:
        a, j = [], []
        for p in fx.parts:
            _build(p, a, j)
        out.append(ExtractionStrings(fx.name, list(fx.extractor_names), "".join(a), "".join(j),
                                     fx.append, getattr(fx, "append_raw", [])))
    return out
