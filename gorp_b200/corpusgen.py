"""Counter-based synthetic corpora (bench / test infrastructure): line i of a corpus is a pure function of (seed, i).

SURVEY.md §8(d): "content of line i = f(seed, i) via a counter-based RNG so shards are identical for any GPU count;
generated on-device". The line grammars are those of gorp_b200/corpus.py (configs #1-#5) restated as small programs for
the generator of csrc/tools/corpusgen.h, which is compiled once into libgorpgen.so and runs on both sides:

    host_text(name, first_line, n)            numpy uint16 text ('\\n'-terminated lines) made on the CPU
    device_text(name, first_line, n, device)  the same units made on the GPU (torch int16 tensor)

tests/test_corpusgen.py checks host == device unit for unit; bench.py shards a corpus over ranks by line index, so the
whole-corpus parity hash (gorp_b200/parityhash.py) is the same number for every GPU count.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import corpus

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "tools", "corpusgen.cu")
HDR = os.path.join(HERE, "csrc", "tools", "corpusgen.h")
LIB = os.path.join(HERE, "libgorpgen.so")

OP = {"END": 0, "LIT": 1, "NUM": 2, "NUMPAD": 3, "TOKEN": 4, "CHOICE": 5, "SKIPIF": 6, "PAD": 7, "IP": 8, "USER": 9, "HEX": 10}
SEEDS = {"simple": 0x5EED0001, "readme": 0x5EED0002, "weblog": 0x5EED0003, "syslog200": 0x5EED0004, "utf16mix": 0x5EED0005}


def build_lib(force=False):
    import fcntl
    from .build import nvcc_path
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        return LIB
    os.makedirs(os.path.join(HERE, "_obj"), exist_ok=True)
    lock = open(os.path.join(HERE, "_obj", ".genlock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        return LIB
    subprocess.check_call([nvcc_path(), "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                           "-Xcompiler", "-fPIC,-Wall", "-shared", "-o", LIB, SRC])
    return LIB


def _p32(p: float) -> int:
    return max(0, min(0xFFFFFFFF, int(round(p * 4294967296.0))))


def _units(s: str):
    return list(np.frombuffer(s.encode("utf-16-le", "surrogatepass"), dtype=np.uint16))


class Program:
    """Builder of a generator program (see csrc/tools/corpusgen.h)."""

    def __init__(self):
        self.strings, self._index, self.ops, self.choice, self.kinds = [], {}, [], [], []
        self.params = [0] * 16
        self.qtable = np.zeros(256, dtype=np.uint16)

    def s(self, text):
        key = text if isinstance(text, str) else tuple(text)
        if key not in self._index:
            u = _units(text) if isinstance(text, str) else list(text)
            self._index[key] = (len(self.strings), len(u))
            self.strings += u
        return self._index[key]

    def lit(self, text):
        off, n = self.s(text)
        if n:
            self.ops.append((OP["LIT"], off, n, 0))

    def num(self, lo, hi):
        self.ops.append((OP["NUM"], lo, hi, 0))

    def numpad(self, lo, hi, width):
        self.ops.append((OP["NUMPAD"], lo, hi, width))

    def hexn(self, lo, hi):
        self.ops.append((OP["HEX"], lo, hi, 0))

    def token(self, alphabet, mn, mx):
        off, n = self.s(alphabet)
        self.ops.append((OP["TOKEN"], off, n | (mn << 16), mx))

    def choose(self, items):
        """items: [(probability, text)]"""
        first = self.choice_table(items)
        self.ops.append((OP["CHOICE"], first, len(items), 0))

    def choice_table(self, items):
        first, cum = len(self.choice), 0.0
        for p, text in items:
            cum += p
            off, n = self.s(text)
            self.choice.append((_p32(cum), off, n))
        return first

    def pad(self, part=0, cap=0, special=False):
        self.ops.append((OP["PAD"], part, cap, 1 if special else 0))

    def ip(self):
        self.ops.append((OP["IP"], 0, 0, 0))

    def user(self):
        self.ops.append((OP["USER"], 0, 0, 0))

    class _Maybe:
        def __init__(self, prog, p_include):
            self.prog, self.p = prog, p_include

        def __enter__(self):
            self.at = len(self.prog.ops)
            self.prog.ops.append(None)

        def __exit__(self, *exc):
            n = len(self.prog.ops) - self.at - 1
            self.prog.ops[self.at] = (OP["SKIPIF"], _p32(1.0 - self.p), n, 0)

    def maybe(self, p_include):
        """with prog.maybe(0.3): ...   -> the enclosed ops are emitted with probability 0.3"""
        return Program._Maybe(self, p_include)

    def kind(self, prob, body):
        self.kinds.append([prob, len(self.ops)])
        body(self)
        self.ops.append((OP["END"], 0, 0, 0))

    def finish(self):
        cum, kinds = 0.0, []
        total = sum(k[0] for k in self.kinds)
        for p, at in self.kinds:
            cum += p / total
            kinds.append((_p32(cum), at))
        kinds[-1] = (0xFFFFFFFF, kinds[-1][1])
        self.a_kinds = np.asarray(kinds, dtype=np.uint32).reshape(-1, 2)
        self.a_ops = np.asarray([[int(v) if v < 2 ** 31 else int(v) - 2 ** 32 for v in o] for o in self.ops], dtype=np.int32).reshape(-1, 4)
        self.a_choice = np.asarray(self.choice if self.choice else [(0, 0, 0)], dtype=np.uint32).reshape(-1, 3)
        self.a_strings = np.asarray(self.strings if self.strings else [0], dtype=np.uint16)
        return self


class _HostProgram(C.Structure):
    _fields_ = [("kinds", C.c_void_p), ("ops", C.c_void_p), ("choice", C.c_void_p), ("strings", C.c_void_p), ("qtable", C.c_void_p),
                ("n_kinds", C.c_uint32), ("n_ops", C.c_uint32), ("n_choice", C.c_uint32), ("n_strings", C.c_uint32),
                ("params", C.c_uint32 * 16)]


_ALNUM38 = "abcdefghijklmnopqrstuvwxyz0123456789-_"
_ALNUM36 = "abcdefghijklmnopqrstuvwxyz0123456789"
_DIGITS = "0123456789"
_MON = ["Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"]
_DOW = ["Mon", "Tue", "Wed", "Thu", "Fri", "Sat", "Sun"]


def _uniform(items):
    return [(1.0 / len(items), x) for x in items]


def _readme_program():
    P = Program()
    verbs = [(0.50, "GET"), (0.30, "PUT"), (0.05, "POST"), (0.05, "DELETE"), (0.05, "HEAD"), (0.05, "PATCH")]

    def ts(p):
        p.num(1, 9)
        p.token(_DIGITS, 8, 9)

    def ms(p):
        p.num(1, 9)
        p.token(_DIGITS, 0, 4)

    def path(p):
        def seg():
            p.lit("/")
            p.token(_ALNUM38, 3, 12)
        seg()
        seg()
        with p.maybe(0.8):
            seg()
            with p.maybe(0.75):
                seg()
                with p.maybe(0.67):
                    seg()
                    with p.maybe(0.5):
                        seg()
        with p.maybe(0.30):
            p.lit("?")
            p.token(_ALNUM38[:26], 3, 8)
            p.lit("=")
            p.token(_ALNUM38, 2, 8)

    def normal(p):
        p.lit("[")
        ts(p)
        p.lit("]: ")
        p.choose(verbs)
        p.lit(" ")
        ms(p)
        p.lit("ms ")
        path(p)

    def readme_style(p):  # the README's own sample line: no brackets, an extra field
        ts(p)
        p.lit(": ")
        p.choose(verbs)
        p.lit(" ")
        ms(p)
        p.lit("ms 200 ")
        path(p)

    def no_ms(p):
        p.lit("[")
        ts(p)
        p.lit("]: ")
        p.choose(verbs)
        p.lit(" ")
        ms(p)
        p.lit(" ")
        path(p)

    P.kind(0.95, normal)
    P.kind(0.025, readme_style)
    P.kind(0.025, no_ms)
    return P.finish()


def _simple_program():
    P = Program()
    tsch = "0123456789-:T+.Z"

    def line(lt=True, status="Accepted", space=True):
        def body(p):
            if lt:
                p.lit("<")
            p.num(0, 191)
            p.lit(">")
            p.token(tsch, 20, 32)
            p.lit(" (" + status + ")" + (" " if space else ""))
        return body

    P.kind(0.50, line())
    P.kind(0.15, line(space=False))
    P.kind(0.15, line(status="Failed"))
    P.kind(0.10, line(lt=False))
    P.kind(0.10, lambda p: p.token("".join(map(chr, range(33, 127))), 1, 48))
    return P.finish()


def _syslog200_program(n_ext=200):
    P = Program()
    spec = corpus.syslog200_spec(n_ext)

    def head(p):
        p.lit("<")
        p.num(0, 191)
        p.lit(">2026-10-17T")
        p.numpad(0, 23, 2)
        p.lit(":")
        p.numpad(0, 59, 2)
        p.lit(":")
        p.numpad(0, 59, 2)
        p.lit(".")
        p.numpad(0, 999, 3)
        p.lit("Z host-")
        p.token(_ALNUM36, 5, 5)
        p.lit(" ")

    def value(p, pat):
        if pat == "%num":
            p.num(0, 999999)
        elif pat == "%word":
            p.token(_ALNUM36, 3, 9)
        elif pat == "%phrase":
            p.lit("/")
            p.token(_ALNUM36, 3, 9)
            p.lit("/")
            p.token(_ALNUM36, 2, 8)
        else:
            p.ip()

    def template(app, parts, tail, variant):
        def body(p):
            head(p)
            if variant == "unknown_app":
                p.lit("daemon_")
                p.num(0, 49)
            else:
                p.lit(app)
            p.lit("[")
            p.num(1, 65535)
            p.lit("]:")
            for j, (words, _, pat) in enumerate(parts):
                p.lit(" " + " ".join(words))
                if variant == "missing_field" and j == len(parts) - 1:
                    return
                p.lit(" ")
                value(p, pat)
            if tail:
                p.lit(" " + " ".join(tail))
            if variant == "junk":
                p.lit(" trailing junk")
        return body

    for variant, share in (("normal", 0.95), ("unknown_app", 0.02), ("missing_field", 0.015), ("junk", 0.015)):
        for app, parts, tail in spec:
            P.kind(share / n_ext, template(app, parts, tail, variant))
    return P.finish()


def _weblog_program(specials: bool):
    P = Program()
    ua = corpus._UA
    msg, amsg = corpus._MSG, corpus._AMSG

    def ts(p):
        p.numpad(1, 28, 2)
        p.lit("/")
        p.choose(_uniform(_MON))
        p.lit("/2026:")
        p.numpad(0, 23, 2)
        p.lit(":")
        p.numpad(0, 59, 2)
        p.lit(":")
        p.numpad(0, 59, 2)
        p.lit(" +0000")

    def access(combined):
        def body(p):
            p.ip()
            p.lit(" - ")
            p.user()
            p.lit(" [")
            ts(p)
            p.lit("] \"")
            p.choose([(.55, "GET"), (.2, "POST"), (.06, "PUT"), (.04, "DELETE"), (.05, "HEAD"), (.05, "OPTIONS"), (.05, "PATCH")])
            p.lit(" /")
            p.token(_ALNUM38, 2, 11)
            with p.maybe(0.75):
                p.lit("/")
                p.token(_ALNUM38, 2, 11)
                with p.maybe(0.67):
                    p.lit("/")
                    p.token(_ALNUM38, 2, 11)
                    with p.maybe(0.5):
                        p.lit("/")
                        p.token(_ALNUM38, 2, 11)
            with p.maybe(0.4):
                p.lit("?q=")
                p.pad(part=1)
            p.lit(" HTTP/1.1\" ")
            p.choose([(3 / 7, "200"), (1 / 7, "301"), (1 / 7, "304"), (1 / 7, "404"), (1 / 7, "500")])
            p.lit(" ")
            with p.maybe(0.1):
                p.lit("-")
            p.num(0, 999999)  # "-123" is still \S+ ($bytes(%phrase)); keeps the op stream free of an else branch
            if combined:
                p.lit(" \"")
                with p.maybe(0.5):
                    p.lit("https://example.com/")
                    p.pad(part=2, special=True)
                p.lit("\" \"")
                p.choose(_uniform(ua))
                p.lit("\"")
        return body

    def ngx_head(p):
        p.lit("2026/")
        p.numpad(1, 12, 2)
        p.lit("/")
        p.numpad(1, 28, 2)
        p.lit(" ")
        p.numpad(0, 23, 2)
        p.lit(":")
        p.numpad(0, 59, 2)
        p.lit(":")
        p.numpad(0, 59, 2)
        p.lit(" [")
        p.choose([(.5, "error"), (.2, "warn"), (.1, "crit"), (.1, "notice"), (.1, "info")])
        p.lit("] ")
        p.num(100, 64999)
        p.lit("#")
        p.num(0, 63)
        p.lit(": ")

    def ngx(client):
        def body(p):
            ngx_head(p)
            if client:
                p.lit("*")
                p.num(1, 999999)
                p.lit(" ")
            p.choose(_uniform(msg))
            with p.maybe(0.3):
                p.lit(" ")
                p.pad(special=True)
            if client:
                p.lit(", client: ")
                p.ip()
                p.lit(", server: example.com, request: \"GET /")
                p.pad(cap=24)
                p.lit(" HTTP/1.1\", host: \"www.example.com\"")
        return body

    def t24(p, short):
        p.choose(_uniform(_DOW))
        p.lit(" ")
        p.choose(_uniform(_MON))
        p.lit(" ")
        p.numpad(1, 28, 2)
        p.lit(" ")
        p.numpad(0, 23, 2)
        p.lit(":")
        p.numpad(0, 59, 2)
        p.lit(":")
        p.numpad(0, 59, 2)
        if not short:
            p.lit(".")
            p.numpad(0, 999999, 6)
        p.lit(" 2026")

    def amessage(p):
        p.choose(_uniform(amsg))
        with p.maybe(0.3):
            p.lit(" ")
            p.pad(special=True)

    def apache(fmt):
        def body(p):
            p.lit("[")
            t24(p, short=fmt in ("22client", "22"))
            if fmt == "24client":
                p.lit("] [core:error] [pid ")
                p.num(100, 64999)
                p.lit(":tid ")
                p.num(10 ** 8, 10 ** 9 - 1)
                p.lit("] [client ")
                p.ip()
                p.lit(":")
                p.num(1024, 65534)
                p.lit("] AH")
                p.numpad(1, 99998, 5)
                p.lit(": ")
            elif fmt == "24":
                p.lit("] [mpm_event:notice] [pid ")
                p.num(100, 64999)
                p.lit(":tid ")
                p.num(10 ** 8, 10 ** 9 - 1)
                p.lit("] AH")
                p.numpad(1, 99998, 5)
                p.lit(": ")
            elif fmt == "22client":
                p.lit("] [error] [client ")
                p.ip()
                p.lit("] ")
            else:
                p.lit("] [notice] ")
            amessage(p)
        return body

    def junk_unquoted(p):
        p.ip()
        p.lit(" - - [")
        ts(p)
        p.lit("] GET /")
        p.pad(cap=20)
        p.lit(" HTTP/1.1 200 12")

    def junk_kv(p):
        p.lit("2026-10-17T13:55:36Z service=")
        p.pad(cap=12)
        p.lit(" level=info msg=\"")
        p.choose(_uniform([m.replace('"', "'") for m in msg]))
        p.lit("\"")

    def junk_trunc(p):
        p.lit("[")
        ts(p)
        p.lit(" [error] [client ")
        p.ip()
        p.lit("] truncated")

    P.kind(0.70 * 0.8, access(True))
    P.kind(0.70 * 0.2, access(False))
    P.kind(0.12 * 0.75, ngx(True))
    P.kind(0.12 * 0.25, ngx(False))
    P.kind(0.08 * 0.35, apache("24client"))
    P.kind(0.08 * 0.20, apache("24"))
    P.kind(0.08 * 0.25, apache("22client"))
    P.kind(0.08 * 0.20, apache("22"))
    P.kind(0.10 * 0.3, junk_unquoted)
    P.kind(0.10 * 0.3, junk_kv)
    P.kind(0.10 * 0.2, junk_trunc)
    P.kind(0.10 * 0.2, lambda p: p.token(_ALNUM38 + " ", 0, 79))
    # the pad field: lognormal lengths clipped to 0..420 units (long referrers / query strings), as 256 quantiles
    rng = np.random.default_rng(7)
    P.qtable = np.sort(np.clip(rng.lognormal(np.log(60), 0.9, size=1 << 16), 0, 420).astype(np.uint16))[127::256].copy()
    off, n = P.s(_ALNUM38)
    P.params[11], P.params[12] = off, n
    if specials:  # config #5: 10 % non-ASCII, 1 % supplementary planes, 0.1 % divergence characters, 0.01 % outliers
        P.params[0] = _p32(0.0001)
        P.params[1] = _p32(0.0011)
        P.params[2] = _p32(0.0111)
        P.params[3] = _p32(0.10)
        P.params[4] = 5000
        P.params[5], P.params[6] = P.choice_table(_uniform(corpus._DIVERGE)), len(corpus._DIVERGE)
        P.params[7], P.params[8] = P.choice_table(_uniform(corpus._SUPPL)), len(corpus._SUPPL)
        P.params[9], P.params[10] = P.choice_table(_uniform(corpus._NONASCII)), len(corpus._NONASCII)
        off, n = P.s("abcdefghij0123456789")
        P.params[13], P.params[14] = off, n
    return P.finish()


_BUILDERS = {"readme": _readme_program, "simple": _simple_program, "syslog200": _syslog200_program,
             "weblog": lambda: _weblog_program(False), "utf16mix": lambda: _weblog_program(True)}
_PROGRAMS = {}
_lib = None


def program(name: str) -> Program:
    if name not in _PROGRAMS:
        _PROGRAMS[name] = _BUILDERS[name]()
    return _PROGRAMS[name]


def _load():
    global _lib
    if _lib is None:
        build_lib()
        _lib = C.CDLL(LIB)
        _lib.cg_lengths.argtypes = [C.POINTER(_HostProgram), C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        _lib.cg_fill.argtypes = [C.POINTER(_HostProgram), C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        _lib.cg_lengths_device.argtypes = [C.POINTER(_HostProgram), C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        _lib.cg_fill_device.argtypes = [C.POINTER(_HostProgram), C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def _host_program(P: Program) -> _HostProgram:
    h = _HostProgram()
    h.kinds, h.ops, h.choice = P.a_kinds.ctypes.data, P.a_ops.ctypes.data, P.a_choice.ctypes.data
    h.strings, h.qtable = P.a_strings.ctypes.data, P.qtable.ctypes.data
    h.n_kinds, h.n_ops, h.n_choice, h.n_strings = len(P.a_kinds), len(P.a_ops), len(P.a_choice), len(P.a_strings)
    for i, v in enumerate(P.params):
        h.params[i] = v
    return h


def host_text(name: str, first_line: int, n: int, seed: int | None = None, threads: int | None = None) -> np.ndarray:
    """Lines [first_line, first_line + n) of corpus `name` as '\\n'-terminated UTF-16 text, generated on the CPU."""
    lib, P = _load(), program(name)
    h = _host_program(P)
    seed = SEEDS[name] if seed is None else seed
    threads = threads or (os.cpu_count() or 1)
    lens = np.empty(n, dtype=np.int32)
    assert lib.cg_lengths(C.byref(h), seed, first_line, n, lens.ctypes.data, threads) == 0
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens.astype(np.int64) + 1, out=off[1:])
    out = np.empty(int(off[-1]), dtype=np.uint16)
    assert lib.cg_fill(C.byref(h), seed, first_line, n, off.ctypes.data, out.ctypes.data, threads) == 0
    return out


def device_text(name: str, first_line: int, n: int, device, seed: int | None = None):
    """The same units generated on `device` (a torch CUDA device): returns an int16 tensor (view it as UTF-16 units)."""
    import torch
    lib, P = _load(), program(name)
    h = _host_program(P)
    seed = SEEDS[name] if seed is None else seed
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream().cuda_stream
        lens = torch.empty(n, dtype=torch.int32, device=device)
        assert lib.cg_lengths_device(C.byref(h), seed, first_line, n, lens.data_ptr(), stream) == 0
        off = torch.zeros(n + 1, dtype=torch.int64, device=device)
        torch.cumsum(lens.to(torch.int64) + 1, dim=0, out=off[1:])
        total = int(off[-1].item())
        del lens
        out = torch.empty(total, dtype=torch.int16, device=device)
        assert lib.cg_fill_device(C.byref(h), seed, first_line, n, off.data_ptr(), out.data_ptr(), stream) == 0
        del off
    return out
