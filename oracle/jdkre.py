"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the capture half of gorp: `java.util.regex.Pattern.compile(s)`
(flags 0) + `Matcher.matches()` + `group(1..n)` as used by
jdkre/JDKRegexpExtractionCooker.java:23 and jdkre/JDKRegexpCookedExtraction.java:36-59.

java.util.regex is JDK code that is absent from /root/reference; its documented
behaviour is restated for the constructs gorp's translators can emit
(util/RegexHelper.java:20-70, :210-237):

  * backtracking, anchored at both ends (`matches()`), preference order:
    alternation left-to-right, greedy quantifiers try one more iteration first,
    lazy ones the reverse; the first full-consumption path wins;
  * CharProperty nodes (negated classes, ranges, `.`, \\D \\S \\W) read a CODE POINT
    (`Character.codePointAt`) and advance 1 or 2 units; BmpCharProperty nodes
    (single literals, small bit classes, \\d \\s \\w) read one UTF-16 unit;
  * `.` excludes \\n \\r U+0085 U+2028 U+2029; \\s = [ \\t\\n\\x0B\\f\\r]; \\w = [a-zA-Z_0-9];
  * group(k) = [start,end) of the LAST participation, null (-1,-1) if none.

Constructs whose JDK meaning differs from brics at the DEFINITION level (\\b, ^, $,
possessive quantifiers, nested classes, &&, (?...) other than (?:, nullable loop
bodies) raise `Unsupported`: the product refuses the same definitions loudly.

"parity unpinned" beyond the reference's own capture vectors
(TST/FullExtractionTest.java, TST/ParametricExtractorTest.java,
TST/ParametricTemplateTest.java): no JVM exists in this environment.
"""
from __future__ import annotations

import sys
import threading

import numpy as np

MAXCP = 0x10FFFF


class JdkSyntaxError(ValueError):
    """Pattern.compile would throw PatternSyntaxException."""


class Unsupported(ValueError):
    """Legal for the reference, but outside the subset the GPU path accepts."""


def _norm(iv):
    iv = sorted((lo, hi) for lo, hi in iv if lo <= hi)
    out = []
    for lo, hi in iv:
        if out and lo <= out[-1][1] + 1:
            if hi > out[-1][1]:
                out[-1] = (out[-1][0], hi)
        else:
            out.append((lo, hi))
    return out


def _complement(iv):
    out, prev = [], 0
    for lo, hi in _norm(iv):
        if lo > prev:
            out.append((prev, lo - 1))
        prev = hi + 1
    if prev <= MAXCP:
        out.append((prev, MAXCP))
    return out


SET_d = [(0x30, 0x39)]
SET_s = _norm([(0x20, 0x20), (0x09, 0x0D)])       # space \t \n \x0B \f \r
SET_w = _norm([(0x61, 0x7A), (0x41, 0x5A), (0x5F, 0x5F), (0x30, 0x39)])
SET_dot = _complement([(0x0A, 0x0A), (0x0D, 0x0D), (0x85, 0x85), (0x2028, 0x2029)])


class Char:
    __slots__ = ("c",)

    def __init__(self, c):
        self.c = c


class Set:
    __slots__ = ("iv", "cpstep", "_los")

    def __init__(self, iv, cpstep):
        self.iv, self.cpstep = _norm(iv), cpstep
        self._los = [a for a, _ in self.iv]

    def contains(self, cp):
        import bisect
        i = bisect.bisect_right(self._los, cp) - 1
        return i >= 0 and self.iv[i][1] >= cp


class Cat:
    __slots__ = ("items",)

    def __init__(self, items):
        self.items = items


class Alt:
    __slots__ = ("items",)

    def __init__(self, items):
        self.items = items


class Group:
    __slots__ = ("idx", "body")

    def __init__(self, idx, body):
        self.idx, self.body = idx, body  # idx None == (?: )


class Rep:
    __slots__ = ("body", "min", "max", "greedy")

    def __init__(self, body, mn, mx, greedy):
        self.body, self.min, self.max, self.greedy = body, mn, mx, greedy


def _nullable(n):
    if isinstance(n, (Char, Set)):
        return False
    if isinstance(n, Cat):
        return all(_nullable(x) for x in n.items)
    if isinstance(n, Alt):
        return any(_nullable(x) for x in n.items)
    if isinstance(n, Group):
        return _nullable(n.body)
    if isinstance(n, Rep):
        return n.min == 0 or _nullable(n.body)
    raise AssertionError


class _Parser:
    def __init__(self, s):
        self.s, self.i, self.ngroups = s, 0, 0
        for ch in s:
            if 0xD800 <= ord(ch) <= 0xDFFF or ord(ch) > 0xFFFF:
                raise Unsupported("surrogate / supplementary character in pattern")

    def peek(self, k=0):
        j = self.i + k
        return self.s[j] if j < len(self.s) else None

    def parse(self):
        e = self.alt()
        if self.i < len(self.s):
            raise JdkSyntaxError("Unmatched closing ')' near index %d" % self.i)
        return e

    def alt(self):
        items = [self.seq()]
        while self.peek() == "|":
            self.i += 1
            items.append(self.seq())
        return items[0] if len(items) == 1 else Alt(items)

    def seq(self):
        items = []
        while True:
            c = self.peek()
            if c is None or c == "|" or c == ")":
                break
            a = self.atom()
            a = self.closure(a)
            items.append(a)
        return Cat(items)

    def atom(self):
        c = self.peek()
        if c == "(":
            self.i += 1
            if self.peek() == "?":
                if self.peek(1) == ":":
                    self.i += 2
                    body = self.alt()
                    idx = None
                else:
                    raise Unsupported("inline construct (?%s" % (self.peek(1) or ""))
            else:
                self.ngroups += 1
                idx = self.ngroups
                body = self.alt()
            if self.peek() != ")":
                raise JdkSyntaxError("Unclosed group near index %d" % self.i)
            self.i += 1
            return Group(idx, body)
        if c == "[":
            self.i += 1
            return self.clazz()
        if c == ".":
            self.i += 1
            return Set(SET_dot, True)
        if c == "\\":
            self.i += 1
            return self.escape_atom()
        if c in "*+?":
            raise JdkSyntaxError("Dangling meta character '%s' near index %d" % (c, self.i))
        if c == "{":
            raise Unsupported("unescaped '{' at start of an atom")
        if c in "^$":
            raise Unsupported("anchor '%s' (literal for brics, anchor for java.util.regex)" % c)
        self.i += 1
        return Char(ord(c))

    def escape_atom(self):
        d = self.peek()
        if d is None:
            raise JdkSyntaxError("Unexpected internal error near index %d" % self.i)
        self.i += 1
        if d == "d":
            return Set(SET_d, False)
        if d == "D":
            return Set(_complement(SET_d), True)
        if d == "s":
            return Set(SET_s, False)
        if d == "S":
            return Set(_complement(SET_s), True)
        if d == "w":
            return Set(SET_w, False)
        if d == "W":
            return Set(_complement(SET_w), True)
        if d == "t":
            return Char(0x09)
        if d == "n":
            return Char(0x0A)
        if d == "r":
            return Char(0x0D)
        if d == "f":
            return Char(0x0C)
        if d == "b":
            raise Unsupported("\\b (backspace for brics, word boundary for java.util.regex)")
        if d.isalpha() or d.isdigit():
            raise Unsupported("escape \\%s" % d)
        return Char(ord(d))

    def _class_escape(self):
        """Returns ('set', iv, is_complement) or ('chr', c)."""
        d = self.peek()
        if d is None:
            raise JdkSyntaxError("Unclosed character class")
        self.i += 1
        if d == "d":
            return ("set", SET_d, False)
        if d == "D":
            return ("set", _complement(SET_d), True)
        if d == "s":
            return ("set", SET_s, False)
        if d == "S":
            return ("set", _complement(SET_s), True)
        if d == "w":
            return ("set", SET_w, False)
        if d == "W":
            return ("set", _complement(SET_w), True)
        if d == "t":
            return ("chr", 0x09)
        if d == "n":
            return ("chr", 0x0A)
        if d == "r":
            return ("chr", 0x0D)
        if d == "f":
            return ("chr", 0x0C)
        if d == "b":
            raise JdkSyntaxError("Illegal/unsupported escape sequence \\b in character class")
        if d.isalpha() or d.isdigit():
            raise Unsupported("escape \\%s" % d)
        return ("chr", ord(d))

    def clazz(self):
        negate = False
        if self.peek() == "^":
            negate = True
            self.i += 1
        iv = []
        have = False
        bits_only = True
        while True:
            c = self.peek()
            if c is None:
                raise JdkSyntaxError("Unclosed character class near index %d" % self.i)
            if c == "]" and have:
                self.i += 1
                break
            if c == "[":
                raise Unsupported("nested character class")
            if c == "&" and self.peek(1) == "&":
                raise Unsupported("character class intersection &&")
            if c == "\\":
                self.i += 1
                r = self._class_escape()
                if r[0] == "set":
                    iv = iv + list(r[1])
                    if r[2]:
                        bits_only = False
                    have = True
                    continue
                lo = r[1]
            else:
                self.i += 1
                lo = ord(c)
            have = True
            if self.peek() == "-":
                e = self.peek(1)
                if e == "[":
                    raise Unsupported("nested character class")
                if e is not None and e != "]":
                    self.i += 1
                    if self.peek() == "\\":
                        self.i += 1
                        r = self._class_escape()
                        if r[0] != "chr":
                            raise JdkSyntaxError("Illegal character range near index %d" % self.i)
                        hi = r[1]
                    else:
                        hi = ord(self.peek())
                        self.i += 1
                    if hi < lo:
                        raise JdkSyntaxError("Illegal character range near index %d" % self.i)
                    iv.append((lo, hi))
                    bits_only = False
                    continue
            iv.append((lo, lo))
            if lo >= 256:
                bits_only = False
        if negate:
            return Set(_complement(iv), True)
        return Set(iv, not bits_only)

    def closure(self, atom):
        c = self.peek()
        if c == "?":
            self.i += 1
            mn, mx = 0, 1
        elif c == "*":
            self.i += 1
            mn, mx = 0, -1
        elif c == "+":
            self.i += 1
            mn, mx = 1, -1
        elif c == "{":
            j = self.i + 1
            st = j
            while j < len(self.s) and self.s[j].isdigit() and self.s[j] in "0123456789":
                j += 1
            if j == st:
                raise JdkSyntaxError("Illegal repetition near index %d" % self.i)
            mn = int(self.s[st:j])
            mx = mn
            if j < len(self.s) and self.s[j] == ",":
                j += 1
                st = j
                while j < len(self.s) and self.s[j] in "0123456789":
                    j += 1
                mx = int(self.s[st:j]) if j > st else -1
                if mx != -1 and mx < mn:
                    raise JdkSyntaxError("Illegal repetition range near index %d" % self.i)
            if j >= len(self.s) or self.s[j] != "}":
                raise JdkSyntaxError("Unclosed counted closure near index %d" % self.i)
            self.i = j + 1
        else:
            return atom
        greedy = True
        n = self.peek()
        if n == "?":
            self.i += 1
            greedy = False
        elif n == "+":
            raise Unsupported("possessive quantifier")
        n = self.peek()
        if n is not None and n in "*+?{":
            raise Unsupported("stacked quantifier")
        if (mx == -1 or mx > 1) and _nullable(atom):
            raise Unsupported("quantified sub-expression can match the empty string")
        return Rep(atom, mn, mx, greedy)


class Pattern:
    def __init__(self, source: str):
        p = _Parser(source)
        self.source = source
        self.ast = p.parse()
        self.ngroups = p.ngroups


def compile(source: str) -> Pattern:  # noqa: A001 - mirrors Pattern.compile
    return Pattern(source)


# --------------------------------------------------------------------------
# Matcher.matches(): AST-walking backtracker with continuations
# --------------------------------------------------------------------------

def to_units(s) -> list:
    """Java String -> UTF-16 code units."""
    if isinstance(s, str):
        return np.frombuffer(s.encode("utf-16-le", "surrogatepass"), dtype="<u2").tolist()
    return [int(x) for x in s]


def _cp_at(u, i, n):
    c = u[i]
    if 0xD800 <= c <= 0xDBFF and i + 1 < n:
        d = u[i + 1]
        if 0xDC00 <= d <= 0xDFFF:
            return 0x10000 + ((c - 0xD800) << 10) + (d - 0xDC00), 2
    return c, 1


def _matches(pat: Pattern, u):
    n = len(u)
    caps = [-1] * (2 * pat.ngroups + 2)

    def step1(node, i):
        """single-width atom: returns new index or -1"""
        if i >= n:
            return -1
        if isinstance(node, Char):
            return i + 1 if u[i] == node.c else -1
        if node.cpstep:
            cp, w = _cp_at(u, i, n)
        else:
            cp, w = u[i], 1
        return i + w if node.contains(cp) else -1

    def m(node, i, k):
        t = type(node)
        if t is Char or t is Set:
            j = step1(node, i)
            return j >= 0 and k(j)
        if t is Cat:
            items = node.items

            def go(idx, j):
                if idx == len(items):
                    return k(j)
                return m(items[idx], j, lambda x: go(idx + 1, x))
            return go(0, i)
        if t is Alt:
            for b in node.items:
                if m(b, i, k):
                    return True
            return False
        if t is Group:
            if node.idx is None:
                return m(node.body, i, k)
            s = 2 * node.idx

            def tail(j):
                old = (caps[s], caps[s + 1])
                caps[s], caps[s + 1] = i, j
                if k(j):
                    return True
                caps[s], caps[s + 1] = old
                return False
            return m(node.body, i, tail)
        if t is Rep:
            body, mn, mx = node.body, node.min, node.max
            if type(body) in (Char, Set):
                if node.greedy:  # Curly.match0: take all, back off one at a time
                    pos = [i]
                    j = i
                    while mx == -1 or len(pos) - 1 < mx:
                        j2 = step1(body, j)
                        if j2 < 0:
                            break
                        pos.append(j2)
                        j = j2
                    cnt = len(pos) - 1
                    if cnt < mn:
                        return False
                    for c in range(cnt, mn - 1, -1):
                        if k(pos[c]):
                            return True
                    return False
                j, cnt = i, 0  # Curly.match1: lazy
                while cnt < mn:
                    j = step1(body, j)
                    if j < 0:
                        return False
                    cnt += 1
                while True:
                    if k(j):
                        return True
                    if mx != -1 and cnt >= mx:
                        return False
                    j = step1(body, j)
                    if j < 0:
                        return False
                    cnt += 1

            def loop(count, j, begin):  # Loop.match / LazyLoop.match
                if j > begin:
                    if count < mn:
                        return m(body, j, lambda x: loop(count + 1, x, j))
                    if node.greedy:
                        if mx == -1 or count < mx:
                            if m(body, j, lambda x: loop(count + 1, x, j)):
                                return True
                    else:
                        if k(j):
                            return True
                        if mx == -1 or count < mx:
                            return m(body, j, lambda x: loop(count + 1, x, j))
                        return False
                return k(j)
            # Prolog -> Loop.matchInit
            if 0 < mn:
                return m(body, i, lambda x: loop(1, x, i))
            if mx == -1 or 0 < mx:
                if node.greedy:
                    if m(body, i, lambda x: loop(1, x, i)):
                        return True
                    return k(i)
                if k(i):
                    return True
                return m(body, i, lambda x: loop(1, x, i))
            return k(i)
        raise AssertionError(t)

    ok = m(pat.ast, 0, lambda j: j == n)  # LastNode with ENDANCHOR
    if not ok:
        return None
    return [(caps[2 * g], caps[2 * g + 1]) for g in range(1, pat.ngroups + 1)]


def matches(pat: Pattern, line):
    """`Matcher.matches()` + `group(1..groupCount())` as (start,end) unit offsets,
    or None when the pattern does not match the whole line."""
    u = to_units(line)
    if len(u) < 200:
        old = sys.getrecursionlimit()
        sys.setrecursionlimit(max(old, 20000))
        try:
            return _matches(pat, u)
        finally:
            sys.setrecursionlimit(old)
    res = []

    def run():
        sys.setrecursionlimit(1000000)
        res.append(_matches(pat, u))
    threading.stack_size(512 * 1024 * 1024)
    t = threading.Thread(target=run)
    t.start()
    t.join()
    threading.stack_size(0)
    return res[0]


# --------------------------------------------------------------------------
# AST -> backtracking program for the C hot loop (oracle/gorp_oracle.c)
# --------------------------------------------------------------------------

OP_CHAR, OP_SET, OP_SPLIT, OP_JMP, OP_SAVE, OP_MATCH = 0, 1, 2, 3, 4, 5
MAX_PROG = 100000


class Program:
    """ops: int32[n,3] (op, a, b); sets: list of (cpstep, intervals)."""

    def __init__(self, pat: Pattern):
        self.ops = []
        self.sets = []
        self._setidx = {}
        self.ngroups = pat.ngroups
        self._emit(pat.ast)
        self.ops.append((OP_MATCH, 0, 0))

    def _set(self, node):
        key = (node.cpstep, tuple(node.iv))
        i = self._setidx.get(key)
        if i is None:
            i = len(self.sets)
            self._setidx[key] = i
            self.sets.append((node.cpstep, list(node.iv)))
        return i

    def _emit(self, n):
        ops = self.ops
        if len(ops) > MAX_PROG:
            raise Unsupported("program too large")
        if isinstance(n, Char):
            ops.append((OP_CHAR, n.c, 0))
        elif isinstance(n, Set):
            ops.append((OP_SET, self._set(n), 0))
        elif isinstance(n, Cat):
            for x in n.items:
                self._emit(x)
        elif isinstance(n, Alt):
            jmps = []
            for k, b in enumerate(n.items):
                if k < len(n.items) - 1:
                    sp = len(ops)
                    ops.append(None)
                    self._emit(b)
                    jmps.append(len(ops))
                    ops.append(None)
                    ops[sp] = (OP_SPLIT, sp + 1, len(ops))
                else:
                    self._emit(b)
            for j in jmps:
                ops[j] = (OP_JMP, len(ops), 0)
        elif isinstance(n, Group):
            if n.idx is None:
                self._emit(n.body)
            else:
                ops.append((OP_SAVE, 2 * n.idx, 0))
                self._emit(n.body)
                ops.append((OP_SAVE, 2 * n.idx + 1, 0))
        elif isinstance(n, Rep):
            for _ in range(n.min):
                self._emit(n.body)
            if n.max == -1:
                sp = len(ops)
                ops.append(None)
                self._emit(n.body)
                ops.append((OP_JMP, sp, 0))
                end = len(ops)
                ops[sp] = (OP_SPLIT, sp + 1, end) if n.greedy else (OP_SPLIT, end, sp + 1)
            else:
                sps = []
                for _ in range(n.max - n.min):
                    sps.append(len(ops))
                    ops.append(None)
                    self._emit(n.body)
                end = len(ops)
                for sp in sps:
                    ops[sp] = (OP_SPLIT, sp + 1, end) if n.greedy else (OP_SPLIT, end, sp + 1)
        else:
            raise AssertionError(n)
