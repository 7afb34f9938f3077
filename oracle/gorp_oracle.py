"""ORACLE — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
arm may import this package; the product (`gorp_b200/`) never does.

`Gorp` here restates the reference driver end to end:
  Gorp.construct  Gorp.java:50-92      -> regex strings, PolyMatcher, one Pattern per extraction
  Gorp.extract    Gorp.java:159-186    -> MISS | MATCH(index, spans) | CAPTURE_FAIL(index)

Per-line outcomes are encoded exactly like the C ABI (include/gorp_cuda.h):
  ext_id >= 0  matched extraction index, spans = (start,end) UTF-16 unit offsets
  ext_id == -1 miss (extract() returns null, Gorp.java:162-164)
  ext_id <= -2 capture failure for extraction -2-ext_id (extract() throws, :173-177)

The Python classes are the readable restatement (small cases); `Gorp.extract_batch`
runs the same algorithm through oracle/gorp_oracle.c for multi-million-line parity
and for bench.py's CPU baseline.

Parity status: pinned for every vector the reference's tests hold for this path
(tests/test_oracle_golden.py); "parity unpinned" for everything else — the reference
is Java and no JVM exists here, and both engines it delegates to (brics automaton
1.11-8, java.util.regex) are third-party code absent from /root/reference.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import brics, frontend, jdkre

MISS = -1

_HERE = os.path.dirname(os.path.abspath(__file__))


class Extraction:
    def __init__(self, index, xs, pattern):
        self.index, self.name = index, xs.name
        self.extractor_names = xs.extractor_names
        self.autom_source, self.regexp_source = xs.autom, xs.jdk
        self.append = xs.append
        self.pattern = pattern


class Gorp:
    def __init__(self, definition: str):
        strings = frontend.definition_to_strings(definition)
        self.extractions = []
        for i, xs in enumerate(strings):
            try:
                pat = jdkre.compile(xs.jdk)  # cooker.cook -> Pattern.compile (Gorp.java:72-77)
            except jdkre.JdkSyntaxError as e:
                raise frontend.DefinitionParseError(
                    "Internal problem: invalid regular expression segment, problem: %s" % e)
            if pat.ngroups != len(xs.extractor_names):
                raise jdkre.Unsupported("capturing groups (%d) != extractors (%d) in %r"
                                        % (pat.ngroups, len(xs.extractor_names), xs.jdk))
            self.extractions.append(Extraction(i, xs, pat))
        try:
            self.matcher = brics.PolyMatcher([x.autom_source for x in self.extractions])
        except brics.BricsSyntaxError as e:
            raise frontend.DefinitionParseError(
                "Internal error: problem with PolyMatcher construction: Invalid regexp, %s" % e)
        self._c = None

    # ---- Gorp.extract (Gorp.java:159-186)
    def extract(self, line):
        u = jdkre.to_units(line)
        idx = self.matcher.match(u)
        if not idx:
            return (MISS, None)
        e = idx[0]
        spans = jdkre.matches(self.extractions[e].pattern, u)
        if spans is None:
            return (-2 - e, None)
        return (e, spans)

    def extract_map(self, line, id_as=None):
        """ExtractionResult.asMap(idAs) (ExtractionResult.java:65-88) or None / raises."""
        u = jdkre.to_units(line)
        e, spans = self.extract(u)
        if e == MISS:
            return None
        if e < 0:
            x = self.extractions[-2 - e]
            raise ExtractionError(
                "Internal error: high-level match for extraction #%d (%s) failed to match generated regexp: %s"
                % (-2 - e, x.name, x.regexp_source))
        x = self.extractions[e]
        out = {}
        if id_as is not None:
            out[id_as] = x.name
        arr = np.asarray(u, dtype="<u2")
        for nm, (a, b) in zip(x.extractor_names, spans):
            out[nm] = None if a < 0 else arr[a:b].tobytes().decode("utf-16-le", "surrogatepass")
        if x.append:
            out.update(x.append)
        return out

    # ---- batch path through the C restatement
    def _cstate(self):
        if self._c is None:
            self._c = _CState(self)
        return self._c

    def extract_batch(self, text: np.ndarray, offsets: np.ndarray, threads: int = 0, max_groups=None):
        """text: uint16[...] ; offsets: int64[n+1] (line i = text[off[i]:off[i+1]), a trailing '\\n'
        belongs to no line when `offsets` came from split_lines).
        Returns (ext_id int32[n], spans int32[n, 2*G]) with G = max group count, -1 padded."""
        return self._cstate().run(text, offsets, threads, max_groups)


class ExtractionError(Exception):
    """Mirrors ExtractionException (Gorp.java:173-177)."""


def split_lines(text: np.ndarray):
    """The CharBuffer line-splitting rule of SURVEY §8(b): split on U+000A only; a final
    line without '\\n' counts; '\\r' is data. Returns (starts int64[n], ends int64[n])."""
    nl = np.flatnonzero(text == 0x0A).astype(np.int64)
    starts = np.concatenate(([0], nl + 1))
    ends = np.concatenate((nl, [len(text)]))
    if len(text) == 0 or (len(nl) and nl[-1] == len(text) - 1):
        starts, ends = starts[:-1], ends[:-1]
    return starts, ends


# --------------------------------------------------------------------------
# C hot loop
# --------------------------------------------------------------------------

def build_c(force=False):
    """Compiles oracle/gorp_oracle.c -> oracle/_build/libgorp_oracle.so (gcc only)."""
    src = os.path.join(_HERE, "gorp_oracle.c")
    outdir = os.path.join(_HERE, "_build")
    out = os.path.join(outdir, "libgorp_oracle.so")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    os.makedirs(outdir, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-pthread", "-o", out, src])
    return out


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c())
        _lib.gorp_oracle_run.restype = ctypes.c_int
    return _lib


class _CState:
    def __init__(self, g: Gorp):
        a = g.matcher.automata
        self.alphabet = np.ascontiguousarray(a.alphabet, dtype=np.int32)
        self.trans = np.ascontiguousarray(a.transitions, dtype=np.int32)
        self.stride = a.stride
        self.accept_first = np.ascontiguousarray(a.accept_first, dtype=np.int32)
        progs = [jdkre.Program(x.pattern) for x in g.extractions]
        self.ngroups = np.asarray([p.ngroups for p in progs], dtype=np.int32)
        # flat program image: ops (int32 triples) with per-extraction offsets; sets flattened
        ops, op_off = [], [0]
        set_hdr, set_iv, set_off = [], [], [0]
        for p in progs:
            base_set = len(set_hdr)
            for (op, x, y) in p.ops:
                if op == jdkre.OP_SET:
                    x += base_set
                ops.append((op, x, y))
            op_off.append(len(ops))
            for cpstep, iv in p.sets:
                set_hdr.append((1 if cpstep else 0, len(set_iv), len(iv)))
                for lo, hi in iv:
                    set_iv.append((lo, hi))
            set_off.append(len(set_hdr))
        self.ops = np.asarray(ops, dtype=np.int32).reshape(-1, 3)
        self.op_off = np.asarray(op_off, dtype=np.int32)
        self.set_hdr = np.asarray(set_hdr, dtype=np.int32).reshape(-1, 3)
        self.set_iv = np.asarray(set_iv, dtype=np.int32).reshape(-1, 2)
        self.maxg = int(self.ngroups.max()) if len(progs) else 0

    def run(self, text, offsets, threads, max_groups=None):
        lib = _load()
        text = np.ascontiguousarray(text, dtype=np.uint16)
        if isinstance(offsets, tuple):
            starts = np.ascontiguousarray(offsets[0], dtype=np.int64)
            ends = np.ascontiguousarray(offsets[1], dtype=np.int64)
        else:
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            starts, ends = offsets[:-1].copy(), offsets[1:].copy()
        n = len(starts)
        G = self.maxg if max_groups is None else max_groups
        ext = np.empty(n, dtype=np.int32)
        spans = np.full((n, 2 * max(G, 1)), -1, dtype=np.int32)
        if threads <= 0:
            threads = os.cpu_count() or 1
        P = ctypes.c_void_p
        rc = lib.gorp_oracle_run(
            P(self.alphabet.ctypes.data), P(self.trans.ctypes.data), ctypes.c_int(self.stride),
            P(self.accept_first.ctypes.data),
            P(self.ops.ctypes.data), P(self.op_off.ctypes.data), P(self.ngroups.ctypes.data),
            P(self.set_hdr.ctypes.data), P(self.set_iv.ctypes.data),
            P(text.ctypes.data), P(starts.ctypes.data), P(ends.ctypes.data), ctypes.c_int64(n),
            P(ext.ctypes.data), P(spans.ctypes.data), ctypes.c_int(2 * max(G, 1)), ctypes.c_int(threads),
            ctypes.c_int(len(self.ngroups)), ctypes.c_int(len(self.set_hdr)))
        if rc != 0:
            raise RuntimeError("gorp_oracle_run failed: %d" % rc)
        return ext, spans[:, :2 * G] if G else spans[:, :0]
