"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the DFA half of gorp:

  * dk.brics.automaton:automaton:1.11-8 `RegExp(s, RegExp.NONE).toAutomaton()` +
    `minimize()` — a THIRD-PARTY dependency that is absent from /root/reference
    (gorp-core/pom.xml:26-30). Its published grammar (RegExp javadoc / source,
    flags = NONE) is restated in `parse()`; everything after parsing is
    language-level automata theory (Thompson NFA -> subset construction ->
    minimisation -> trim), so any correct construction gives the same minimal DFA.
    Call sites it is anchored on: autom/PolyMatcher.java:76-77,
    autom/Automata.java:60,145, autom/PolyState.java:59,68,
    autom/DkBricsAutomatonAccess.java:29-43.
  * gorp's own product construction, restated step by step:
    autom/Automata.java:45-55 (alphabet), :57-124 (construct, BFS numbering),
    :133-139 (step/accept), :150-166 (pointsUnion); autom/PolyState.java:46-90;
    autom/PolyMatcher.java:123-133 (match).

Pinned by the reference's own vectors: TST/autom/MultiPatternTest.java:12-27 and
TST/PolyMatchTest.java (see tests/test_oracle_golden.py).
"""
from __future__ import annotations

import bisect
from collections import deque

import numpy as np

MAXC = 0xFFFF


# --------------------------------------------------------------------------
# interval sets over UTF-16 code units
# --------------------------------------------------------------------------

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


def _complement(iv, maxc=MAXC):
    out, prev = [], 0
    for lo, hi in _norm(iv):
        if lo > prev:
            out.append((prev, lo - 1))
        prev = hi + 1
    if prev <= maxc:
        out.append((prev, maxc))
    return out


# --------------------------------------------------------------------------
# RegExp parser (flags = NONE): union > concat > repeat > charclass > simple
# --------------------------------------------------------------------------

class BricsSyntaxError(ValueError):
    pass


class _Parser:
    def __init__(self, s):
        self.b, self.pos = s, 0

    def more(self):
        return self.pos < len(self.b)

    def peek(self, chars):
        return self.more() and self.b[self.pos] in chars

    def match(self, c):
        if self.pos >= len(self.b):
            return False
        if self.b[self.pos] == c:
            self.pos += 1
            return True
        return False

    def next(self):
        if not self.more():
            raise BricsSyntaxError("unexpected end-of-string")
        c = self.b[self.pos]
        self.pos += 1
        return c

    def parse(self):
        if len(self.b) == 0:
            return ("string", "")
        e = self.union()
        if self.pos < len(self.b):
            raise BricsSyntaxError("end-of-string expected at position %d" % self.pos)
        return e

    def union(self):
        e = self.concat()  # parseInterExp == parseConcatExp when INTERSECTION is off
        if self.match("|"):
            e = ("union", e, self.union())
        return e

    def concat(self):
        e = self.repeat()
        if self.more() and not self.peek(")|"):
            e = ("concat", e, self.concat())
        return e

    def repeat(self):
        e = self.charclass_exp()  # COMPLEMENT is off
        while self.peek("?*+{"):
            if self.match("?"):
                e = ("repeat", e, 0, 1)
            elif self.match("*"):
                e = ("repeat", e, 0, -1)
            elif self.match("+"):
                e = ("repeat", e, 1, -1)
            elif self.match("{"):
                start = self.pos
                while self.peek("0123456789"):
                    self.next()
                if start == self.pos:
                    raise BricsSyntaxError("integer expected at position %d" % self.pos)
                n = int(self.b[start:self.pos])
                m = -1
                if self.match(","):
                    start = self.pos
                    while self.peek("0123456789"):
                        self.next()
                    if start != self.pos:
                        m = int(self.b[start:self.pos])
                else:
                    m = n
                if not self.match("}"):
                    raise BricsSyntaxError("expected '}' at position %d" % self.pos)
                e = ("repeat", e, n, m)
        return e

    def charclass_exp(self):
        if self.match("["):
            negate = self.match("^")
            iv = self.charclasses()
            if negate:
                iv = _complement(iv)
            if not self.match("]"):
                raise BricsSyntaxError("expected ']' at position %d" % self.pos)
            return ("set", _norm(iv))
        return self.simple()

    def charclasses(self):
        iv = self.charclass()
        while self.more() and not self.peek("]"):
            iv = iv + self.charclass()
        return iv

    def charclass(self):
        c = self.char_exp()
        if self.match("-"):
            if self.peek("]"):
                return [(ord(c), ord(c)), (ord("-"), ord("-"))]
            d = self.char_exp()
            return [(ord(c), ord(d))]  # reversed range == empty (Automaton.makeCharRange)
        return [(ord(c), ord(c))]

    def simple(self):
        if self.match("."):
            return ("set", [(0, MAXC)])
        if self.match('"'):
            start = self.pos
            while self.more() and not self.peek('"'):
                self.next()
            if not self.match('"'):
                raise BricsSyntaxError("expected '\"' at position %d" % self.pos)
            return ("string", self.b[start:self.pos - 1])
        if self.match("("):
            if self.match(")"):
                return ("string", "")
            e = self.union()
            if not self.match(")"):
                raise BricsSyntaxError("expected ')' at position %d" % self.pos)
            return e
        c = self.char_exp()
        return ("set", [(ord(c), ord(c))])

    def char_exp(self):
        self.match("\\")
        return self.next()


def parse(s: str):
    return _Parser(s).parse()


# --------------------------------------------------------------------------
# AST -> NFA (Thompson) -> DFA -> minimal trimmed DFA
# --------------------------------------------------------------------------

class _NFA:
    def __init__(self):
        self.eps = []    # state -> list of states
        self.edges = []  # state -> list of (intervals, dest)

    def new(self):
        self.eps.append([])
        self.edges.append([])
        return len(self.eps) - 1


MAX_UNROLL_STATES = 200000


def _build(nfa, e):
    """Returns (start, end) fragment."""
    k = e[0]
    if k == "string":
        s = nfa.new()
        cur = s
        for ch in e[1]:
            t = nfa.new()
            nfa.edges[cur].append(([(ord(ch), ord(ch))], t))
            cur = t
        return s, cur
    if k == "set":
        s, t = nfa.new(), nfa.new()
        if e[1]:
            nfa.edges[s].append((e[1], t))
        return s, t
    if k == "union":
        s, t = nfa.new(), nfa.new()
        for sub in (e[1], e[2]):
            a, b = _build(nfa, sub)
            nfa.eps[s].append(a)
            nfa.eps[b].append(t)
        return s, t
    if k == "concat":
        a, b = _build(nfa, e[1])
        c, d = _build(nfa, e[2])
        nfa.eps[b].append(c)
        return a, d
    if k == "repeat":
        sub, n, m = e[1], e[2], e[3]
        s = nfa.new()
        cur = s
        if m != -1 and n > m:  # Automaton.repeat(min,max): min > max -> empty language
            return s, nfa.new()
        for _ in range(n):
            a, b = _build(nfa, sub)
            nfa.eps[cur].append(a)
            cur = b
            if len(nfa.eps) > MAX_UNROLL_STATES:
                raise BricsSyntaxError("repeat unrolls too far")
        if m == -1:
            a, b = _build(nfa, sub)
            loop = nfa.new()
            nfa.eps[cur].append(loop)
            nfa.eps[loop].append(a)
            nfa.eps[b].append(loop)
            return s, loop
        end = nfa.new()
        nfa.eps[cur].append(end)
        for _ in range(m - n):
            a, b = _build(nfa, sub)
            nfa.eps[cur].append(a)
            cur = b
            nfa.eps[cur].append(end)
            if len(nfa.eps) > MAX_UNROLL_STATES:
                raise BricsSyntaxError("repeat unrolls too far")
        return s, end
    raise AssertionError(k)


class MinDFA:
    """Minimal, trimmed DFA with interval transitions (what brics holds after
    `minimize()`); state 0 is initial; `trans[s]` = sorted [(lo, hi, dest)] with
    maximal intervals per destination (Automaton.reduce())."""

    def __init__(self, trans, accept):
        self.trans, self.accept = trans, accept
        self._los = [[t[0] for t in row] for row in trans]

    def step(self, s, c):  # State.step(char): None when no interval covers c
        row = self.trans[s]
        i = bisect.bisect_right(self._los[s], c) - 1
        if i >= 0 and row[i][1] >= c:
            return row[i][2]
        return -1

    def start_points(self):  # Automaton.getStartPoints()
        pts = {0}
        for row in self.trans:
            for lo, hi, _ in row:
                pts.add(lo)
                if hi < MAXC:
                    pts.add(hi + 1)
        return sorted(pts)


def to_min_dfa(regex: str) -> MinDFA:
    ast = parse(regex)
    nfa = _NFA()
    start, end = _build(nfa, ast)
    # alphabet partition of this regex
    cuts = {0}
    for edges in nfa.edges:
        for iv, _ in edges:
            for lo, hi in iv:
                cuts.add(lo)
                if hi < MAXC:
                    cuts.add(hi + 1)
    cuts = sorted(cuts)
    ncls = len(cuts)

    def closure(states):
        seen = set(states)
        stack = list(states)
        while stack:
            s = stack.pop()
            for t in nfa.eps[s]:
                if t not in seen:
                    seen.add(t)
                    stack.append(t)
        return frozenset(seen)

    # per NFA edge, which classes it covers
    edge_cls = {}
    for s, edges in enumerate(nfa.edges):
        for iv, d in edges:
            cl = []
            for lo, hi in iv:
                a = bisect.bisect_left(cuts, lo)
                b = bisect.bisect_right(cuts, hi)
                cl.extend(range(a, b))
            edge_cls.setdefault(s, []).append((cl, d))

    init = closure([start])
    ids = {init: 0}
    order = [init]
    table = []
    i = 0
    while i < len(order):
        cur = order[i]
        i += 1
        moves = [set() for _ in range(ncls)]
        for s in cur:
            for cl, d in edge_cls.get(s, ()):
                for c in cl:
                    moves[c].add(d)
        row = []
        for c in range(ncls):
            if not moves[c]:
                row.append(-1)
                continue
            nxt = closure(moves[c])
            j = ids.get(nxt)
            if j is None:
                j = len(order)
                ids[nxt] = j
                order.append(nxt)
            row.append(j)
        table.append(row)
    n = len(order)
    acc = [end in st for st in order]

    # Moore minimisation on the total DFA (dead state = n)
    tot = [row[:] for row in table] + [[n] * ncls]
    for row in tot:
        for c in range(ncls):
            if row[c] == -1:
                row[c] = n
    part = [1 if (s < n and acc[s]) else 0 for s in range(n + 1)]
    while True:
        sig = {}
        newpart = []
        for s in range(n + 1):
            key = (part[s], tuple(part[tot[s][c]] for c in range(ncls)))
            newpart.append(sig.setdefault(key, len(sig)))
        if len(sig) == len(set(part)):
            part = newpart
            break
        part = newpart
    nblocks = len(set(part))
    rep = {}
    for s in range(n + 1):
        rep.setdefault(part[s], s)
    btrans = {b: [part[tot[rep[b]][c]] for c in range(ncls)] for b in rep}
    bacc = {b: (rep[b] < n and acc[rep[b]]) for b in rep}
    # live blocks: can reach an accepting block
    live = {b for b in rep if bacc[b]}
    changed = True
    while changed:
        changed = False
        for b in rep:
            if b not in live and any(t in live for t in btrans[b]):
                live.add(b)
                changed = True
    # renumber reachable live blocks from the initial block (BFS)
    b0 = part[0]
    num = {b0: 0}
    q = deque([b0])
    olist = [b0]
    while q:
        b = q.popleft()
        for t in btrans[b]:
            if t in live and t not in num:
                num[t] = len(olist)
                olist.append(t)
                q.append(t)
    trans, accept = [], []
    for b in olist:
        row = []
        if b in live or b == b0:
            c = 0
            while c < ncls:
                t = btrans[b][c]
                if t in live:
                    c2 = c
                    while c2 + 1 < ncls and btrans[b][c2 + 1] == t:
                        c2 += 1
                    hi = (cuts[c2 + 1] - 1) if c2 + 1 < ncls else MAXC
                    row.append((cuts[c], hi, num[t]))
                    c = c2 + 1
                else:
                    c += 1
        trans.append(row)
        accept.append(bool(bacc[b]))
    del nblocks
    return MinDFA(trans, accept)


# --------------------------------------------------------------------------
# Automata.java — the product construction, and PolyMatcher.match
# --------------------------------------------------------------------------

class Automata:
    """Field-for-field restatement of autom/Automata.java:23-31:
    `accept` (int[][]), `stride`, `transitions` (int[]), `alphabet` (int[65536])."""

    def __init__(self, dfas):
        self.n_regex = len(dfas)
        pts = set()
        for d in dfas:  # pointsUnion (:150-166)
            pts.update(d.start_points())
        points = sorted(pts)
        self.points = points
        plen = len(points)
        # dense per-component tables over the union points; last row == null state
        offs, rows, accs = [], [], []
        total = 0
        for d in dfas:
            offs.append(total)
            for s in range(len(d.trans)):
                rows.append([d.step(s, p) for p in points])
                accs.append(d.accept[s])
            total += len(d.trans)
        null = total
        big = np.full((total + 1, plen), -1, dtype=np.int64)
        for i, d in enumerate(dfas):
            o = offs[i]
            for s in range(len(d.trans)):
                r = np.asarray(rows[o + s], dtype=np.int64)
                big[o + s] = np.where(r >= 0, r + o, null)
        big[null] = null
        acc_flat = np.zeros(total + 1, dtype=bool)
        acc_flat[:total] = accs
        init = np.asarray(offs, dtype=np.int64)  # every component's initial state is 0
        index = {init.tobytes(): 0}
        queue = deque([init])
        trans_rows = []
        accept = [np.nonzero(acc_flat[init])[0].tolist()]
        while queue:  # BFS (:67-100): classes ascending, ids in discovery order
            v = queue.popleft()
            nxt = big[v]                      # [N, plen]
            isnull = (nxt == null).all(axis=0)
            row = np.empty(plen, dtype=np.int32)
            for c in range(plen):
                if isnull[c]:
                    row[c] = -1
                    continue
                col = np.ascontiguousarray(nxt[:, c])
                key = col.tobytes()
                j = index.get(key)
                if j is None:
                    j = len(index)
                    index[key] = j
                    queue.append(col)
                    accept.append(np.nonzero(acc_flat[col])[0].tolist())
                row[c] = j
            trans_rows.append(row)
        self.n_states = len(trans_rows)
        self.stride = plen
        self.transitions = np.concatenate(trans_rows).astype(np.int32)
        self.accept = accept
        # alphabet(points) (:45-55)
        self.alphabet = (np.searchsorted(np.asarray(points), np.arange(65536), side="right") - 1).astype(np.int32)
        self.accept_first = np.asarray([a[0] if a else -1 for a in accept], dtype=np.int32)

    def step(self, state, c):  # :133-135
        return int(self.transitions[state * self.stride + self.alphabet[c]])


class PolyMatcher:
    def __init__(self, patterns):
        self.dfas = [to_min_dfa(p) for p in patterns]
        self.automata = Automata(self.dfas)

    def match(self, units):  # PolyMatcher.java:123-133
        p = 0
        a = self.automata
        for c in units:
            p = a.step(p, c)
            if p == -1:
                return []
        return list(a.accept[p])
