"""Synthetic log corpora for the BASELINE.json configs, vectorised numpy, seeded.

Every generator returns '\\n'-separated UTF-16 text as a uint16 array (each line ends in '\\n').
Definitions for the configs live here too so tests, bench.py and smoke() share them.
"""
from __future__ import annotations

import numpy as np

# config #1: /samples/simple.grp of the reference (trailing space after the extractor is significant)
SIMPLE_GRP = (
    "pattern %ws \\s+\n"
    "pattern %optws \\s*\n"
    "pattern %word \\w+\n"
    "pattern %phrase \\S+\n"
    "pattern %num \\d+\n"
    "pattern %ts %phrase\n"
    "pattern %ip %phrase\n"
    "pattern %any .*\n"
    "template @base <%num>$eventTimeStamp(%ts)\n"
    "extract sampleMatch {\n"
    "  template @base ($authStatus(Accepted)) \n"
    "}\n"
)

# config #2: README.md:115-134 of the reference, verbatim
README_DEF = (
    "pattern %num \\d+\n"
    "pattern %word \\w+\n"
    "pattern %phrase \\S+\n"
    "\n"
    "extract PutRequest {\n"
    "   # It's ok to: (a) extract constant value; (b) concatenate physical lines with backslash\n"
    "   template [$timestamp(%num)]: $verb(PUT) $timeTakenInMsec(%num)ms\\\n"
    " $path(%phrase)\n"
    "   append { \"marker\" : \"EXTRACTED\" }\n"
    "}\n"
    "extract GetRequest {\n"
    "   template [$timestamp(%num)]: $verb(GET) $timeTakenInMsec(%num)ms\\\n"
    " $path(%phrase)\n"
    "   append { \"marker\" : \"EXTRACTED\" }\n"
    "}\n"
    "extract OtherRequest {\n"
    "   template [$timestamp(%num)]: $verb(%word) $timeTakenInMsec(%num)ms\\\n"
    " $path(%phrase)\n"
    "   append { \"marker\" : \"EXTRACTED\" }\n"
    "}\n"
)

_ALNUM = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789-_", dtype=np.uint8)
_TSCH = np.frombuffer(b"0123456789-:T+.Z", dtype=np.uint8)


def _assemble(n, comps):
    """comps: list of (matrix uint8/uint16 [n, w], lengths int [n]). Concatenates per line and appends '\\n'."""
    lens = [np.asarray(l, dtype=np.int64) for _, l in comps]
    total = sum(lens) + 1
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(total, out=off[1:])
    out = np.empty(int(off[-1]), dtype=np.uint16)
    out[off[1:] - 1] = 0x0A
    start = off[:-1].copy()
    for (mat, _), ln in zip(comps, lens):
        w = mat.shape[1]
        if w:
            cols = np.arange(w, dtype=np.int64)[None, :]
            mask = cols < ln[:, None]
            pos = start[:, None] + cols
            out[pos[mask]] = mat[mask]
        start += ln
    return out


def _const(n, s: bytes, present=None):
    mat = np.broadcast_to(np.frombuffer(s, dtype=np.uint8), (n, len(s)))
    ln = np.full(n, len(s), dtype=np.int64) if present is None else np.where(present, len(s), 0)
    return mat, ln


def _digits(rng, n, lo, hi):
    mat = rng.integers(48, 58, size=(n, hi), dtype=np.uint8)
    mat[:, 0] = rng.integers(49, 58, size=n, dtype=np.uint8)
    return mat, rng.integers(lo, hi + 1, size=n)


def _choice(rng, n, words, probs):
    w = max(len(x) for x in words)
    table = np.zeros((len(words), w), dtype=np.uint8)
    for i, x in enumerate(words):
        table[i, :len(x)] = np.frombuffer(x, dtype=np.uint8)
    idx = rng.choice(len(words), size=n, p=probs)
    return table[idx], np.asarray([len(x) for x in words], dtype=np.int64)[idx], idx


def readme_corpus(n_lines: int, seed: int = 0x5EED0002) -> np.ndarray:
    """Config #2: `[TS]: VERB MSms PATH`; GET 50 %, PUT 30 %, other verbs 15 %, 5 % non-matching
    (README-style line without brackets and with an extra field, or `ms` missing)."""
    rng = np.random.default_rng(seed)
    n = n_lines
    verbs = [b"GET", b"PUT", b"POST", b"DELETE", b"HEAD", b"PATCH"]
    vm, vl, _ = _choice(rng, n, verbs, [0.50, 0.30, 0.05, 0.05, 0.05, 0.05])
    bad = rng.random(n) < 0.05
    bad_kind = rng.random(n) < 0.5            # True: README-style, False: 'ms' missing
    readme_style = bad & bad_kind
    no_ms = bad & ~bad_kind
    comps = [_const(n, b"[", ~readme_style), _digits(rng, n, 9, 10), _const(n, b"]", ~readme_style), _const(n, b": "),
             (vm, vl), _const(n, b" "), _digits(rng, n, 1, 5), _const(n, b"ms", ~no_ms), _const(n, b" 200", readme_style),
             _const(n, b" ")]
    nseg = rng.integers(2, 7, size=n)
    for s in range(6):
        comps.append(_const(n, b"/", nseg > s))
        seg = _ALNUM[rng.integers(0, len(_ALNUM), size=(n, 12))]
        comps.append((seg, np.where(nseg > s, rng.integers(3, 13, size=n), 0)))
    q = rng.random(n) < 0.30
    comps.append(_const(n, b"?", q))
    comps.append((_ALNUM[rng.integers(0, 26, size=(n, 8))], np.where(q, rng.integers(3, 9, size=n), 0)))
    comps.append(_const(n, b"=", q))
    comps.append((_ALNUM[rng.integers(0, len(_ALNUM), size=(n, 8))], np.where(q, rng.integers(2, 9, size=n), 0)))
    return _assemble(n, comps)


def simple_corpus(n_lines: int, seed: int = 0x5EED0001) -> np.ndarray:
    """Config #1: `<PRI>TS (Accepted) ` (trailing space mandatory). 50 % matching; non-matching =
    no trailing space 15 %, `(Failed) ` 15 %, missing '<' 10 %, random printable 10 %."""
    rng = np.random.default_rng(seed)
    n = n_lines
    r = rng.random(n)
    no_space, failed, no_lt, junk = (r >= 0.50) & (r < 0.65), (r >= 0.65) & (r < 0.80), (r >= 0.80) & (r < 0.90), r >= 0.90
    pri = rng.integers(0, 192, size=n)
    pm = np.zeros((n, 3), dtype=np.uint8)
    pl = np.where(pri >= 100, 3, np.where(pri >= 10, 2, 1))
    s = np.char.encode(pri.astype(str), "ascii")
    pm = np.frombuffer(np.char.ljust(s, 3).tobytes(), dtype=np.uint8).reshape(n, 3)
    ts = _TSCH[rng.integers(0, len(_TSCH), size=(n, 32))]
    tl = rng.integers(20, 33, size=n)
    status, sl, _ = _choice(rng, n, [b"Accepted", b"Failed"], [0.5, 0.5])
    sm = np.where(failed[:, None], np.frombuffer(b"Failed  ", dtype=np.uint8)[None, :], np.frombuffer(b"Accepted", dtype=np.uint8)[None, :])
    sl = np.where(failed, 6, 8)
    del status
    good = ~junk
    comps = [_const(n, b"<", good & ~no_lt), (pm, np.where(good, pl, 0)), _const(n, b">", good), (ts, np.where(good, tl, 0)),
             _const(n, b" (", good), (sm, np.where(good, sl, 0)), _const(n, b")", good), _const(n, b" ", good & ~no_space),
             (rng.integers(33, 127, size=(n, 48), dtype=np.uint8), np.where(junk, rng.integers(1, 49, size=n), 0))]
    return _assemble(n, comps)


CONFIGS = {
    "simple": (SIMPLE_GRP, simple_corpus),
    "readme": (README_DEF, readme_corpus),
}
