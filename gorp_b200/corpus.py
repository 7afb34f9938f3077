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


# ---------------------------------------------------------------------------------------------------------------
# config #3: ~20-extraction nginx / Apache access + error definition written with parametric templates
# (README.md:157-181 idiom of the reference), mixed line lengths.
_ACCESS_VERBS = ["GET", "POST", "PUT", "DELETE", "HEAD"]


def _weblog_definition() -> str:
    d = [
        "# nginx / Apache access and error logs",
        "pattern %num \\d+",
        "pattern %word \\w+",
        "pattern %phrase \\S+",
        "pattern %ip [0-9a-fA-F\\.:]+",
        "pattern %ts [^\\]]+",
        "pattern %qstr [^\\\"]*",
        "pattern %proto HTTP/[0-9\\.]+",
        "pattern %date \\d{4}/\\d{2}/\\d{2}",
        "pattern %time \\d{2}:\\d{2}:\\d{2}",
        "pattern %ahcode AH\\d+",
        "pattern %any .*",
        "template @bracketed() [$1(@2)]",
        "template @quoted() \"$1(@2)\"",
        "template @tsdef %ts",
        "template @qdef %qstr",
        "template @clientHead $client(%ip) $ident(%phrase) $user(%phrase) @bracketed($time,@tsdef)",
        "template @statusBytes $status(%num) $bytes(%phrase)",
        "template @refAgent @quoted($referrer,@qdef) @quoted($agent,@qdef)",
        "template @ngxHead $date(%date) $time(%time)",
        "template @ngxIds $pid(%num)#$tid(%num):",
        "template @ngxTail , client: $client(%ip), server: $server(%phrase), request: @quoted($request,@qdef),\\",
        " host: @quoted($host,@qdef)",
    ]
    for v in _ACCESS_VERBS:
        d += ["extract Combined%s {" % v.capitalize(),
              "  template @clientHead \"$verb(%s) $path(%%phrase) $proto(%%proto)\" @statusBytes @refAgent" % v,
              "  append { \"format\" : \"combined\", \"verb_class\" : \"%s\" }" % v.lower(),
              "}"]
    d += ["extract CombinedOther {",
          "  template @clientHead \"$verb(%word) $path(%phrase) $proto(%proto)\" @statusBytes @refAgent",
          "  append { \"format\" : \"combined\" }",
          "}"]
    for v in ("GET", "POST"):
        d += ["extract Common%s {" % v.capitalize(),
              "  template @clientHead \"$verb(%s) $path(%%phrase) $proto(%%proto)\" @statusBytes" % v,
              "  append { \"format\" : \"common\" }",
              "}"]
    d += ["extract CommonOther {",
          "  template @clientHead \"$verb(%word) $path(%phrase) $proto(%proto)\" @statusBytes",
          "}"]
    for lv in ("error", "warn", "crit"):
        d += ["extract NginxClient%s {" % lv.capitalize(),
              "  template @ngxHead [$level(%s)] @ngxIds *$cid(%%num) $message(%%any)@ngxTail" % lv,
              "  append { \"source\" : \"nginx\", \"severity\" : \"%s\" }" % lv,
              "}"]
    d += ["extract NginxClientOther {",
          "  template @ngxHead [$level(%word)] @ngxIds *$cid(%num) $message(%any)@ngxTail",
          "}",
          "extract NginxPlain {",
          "  template @ngxHead [$level(%word)] @ngxIds $message(%any)",
          "}",
          "extract Apache24Client {",
          "  template @bracketed($time,@tsdef) [$module(%word):$level(%word)] [pid $pid(%num):tid $tid(%num)]\\",
          " [client $client(%phrase)] $code(%ahcode): $message(%any)",
          "}",
          "extract Apache24 {",
          "  template @bracketed($time,@tsdef) [$module(%word):$level(%word)] [pid $pid(%num):tid $tid(%num)]\\",
          " $code(%ahcode): $message(%any)",
          "}",
          "extract Apache22Client {",
          "  template @bracketed($time,@tsdef) [$level(%word)] [client $client(%ip)] $message(%any)",
          "}",
          "extract Apache22 {",
          "  template @bracketed($time,@tsdef) [$level(%word)] $message(%any)",
          "}"]
    return "\n".join(d) + "\n"


WEBLOG_DEF = _weblog_definition()

_UA = ["Mozilla/5.0 (X11; Linux x86_64) AppleWebKit/537.36 (KHTML, like Gecko) Chrome/126.0.0.0 Safari/537.36",
       "curl/8.5.0", "Mozilla/5.0 (Macintosh; Intel Mac OS X 14_5) AppleWebKit/605.1.15 (KHTML, like Gecko) Version/17.5 Safari/605.1.15",
       "python-requests/2.32.3", "Googlebot/2.1 (+http://www.google.com/bot.html)", "-",
       "Mozilla/5.0 (Windows NT 10.0; Win64; x64; rv:127.0) Gecko/20100101 Firefox/127.0"]
_MON = ["Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"]
_DOW = ["Mon", "Tue", "Wed", "Thu", "Fri", "Sat", "Sun"]
_MSG = ["upstream timed out (110: Connection timed out) while reading response header from upstream",
        "open() \"/var/www/html/favicon.ico\" failed (2: No such file or directory)",
        "client intended to send too large body: 10485761 bytes", "SSL_do_handshake() failed",
        "connect() failed (111: Connection refused) while connecting to upstream",
        "recv() failed (104: Connection reset by peer)", "limiting requests, excess: 5.320 by zone \"api\""]
_AMSG = ["File does not exist: /var/www/html/robots.txt", "client denied by server configuration: /srv/private",
         "script not found or unable to stat: /usr/lib/cgi-bin/php", "request failed: error reading the headers",
         "caught SIGTERM, shutting down", "Invalid URI in request GET /%% HTTP/1.1"]


def _pad_token(rng, alphabet, n):
    return "".join(alphabet[i] for i in rng.integers(0, len(alphabet), size=n))


def weblog_lines(n_lines: int, seed: int = 0x5EED0003, special=None):
    """Config #3 lines as Python strings. Access lines (70 %), nginx error (12 %), Apache error (8 %), 10 % that match
    nothing. Lengths roughly lognormal, clipped to 60..600 units (long referrers / query strings / agents).
    `special(rng, i, fields)` may rewrite the variable fields of line i (config #5 uses it)."""
    rng = np.random.default_rng(seed)
    alnum = "abcdefghijklmnopqrstuvwxyz0123456789-_"
    out = []
    kind = rng.random(n_lines)
    extra = np.clip(rng.lognormal(np.log(60), 0.9, size=n_lines), 0, 420).astype(np.int64)  # padding units for long fields
    for i in range(n_lines):
        k = kind[i]
        ip = "%d.%d.%d.%d" % tuple(rng.integers(1, 255, size=4)) if rng.random() < 0.9 else "2001:db8::%x:%x" % tuple(rng.integers(1, 65535, size=2))
        f = {"ip": ip, "pad": _pad_token(rng, alnum, int(extra[i])), "user": "-" if rng.random() < 0.8 else _pad_token(rng, alnum, 6)}
        if special is not None:
            special(rng, i, f)
        ts = "%02d/%s/2026:%02d:%02d:%02d +0000" % (rng.integers(1, 29), _MON[rng.integers(0, 12)], rng.integers(0, 24), rng.integers(0, 60), rng.integers(0, 60))
        if k < 0.70:
            verb = ["GET", "POST", "PUT", "DELETE", "HEAD", "OPTIONS", "PATCH"][rng.choice(7, p=[.55, .2, .06, .04, .05, .05, .05])]
            path = "/" + "/".join(_pad_token(rng, alnum, int(rng.integers(2, 12))) for _ in range(rng.integers(1, 5)))
            if rng.random() < 0.4:
                path += "?q=" + f["pad"][: len(f["pad"]) // 2]
            line = '%s - %s [%s] "%s %s HTTP/1.1" %d %s' % (f["ip"], f["user"], ts, verb, path, [200, 200, 200, 301, 304, 404, 500][rng.integers(0, 7)],
                                                            "-" if rng.random() < 0.1 else str(rng.integers(0, 10**6)))
            if rng.random() < 0.8:  # combined; else common
                line += ' "%s" "%s"' % ("-" if rng.random() < 0.5 else "https://example.com/" + f["pad"][len(f["pad"]) // 2:], _UA[rng.integers(0, len(_UA))])
        elif k < 0.82:
            lv = ["error", "warn", "crit", "notice", "info"][rng.choice(5, p=[.5, .2, .1, .1, .1])]
            head = "2026/%02d/%02d %02d:%02d:%02d [%s] %d#%d: " % (rng.integers(1, 13), rng.integers(1, 29), rng.integers(0, 24), rng.integers(0, 60), rng.integers(0, 60), lv,
                                                                  rng.integers(100, 65000), rng.integers(0, 64))
            msg = _MSG[rng.integers(0, len(_MSG))] + (" " + f["pad"] if rng.random() < 0.3 else "")
            if rng.random() < 0.75:
                line = head + "*%d %s, client: %s, server: %s, request: \"GET /%s HTTP/1.1\", host: \"%s\"" % (
                    rng.integers(1, 10**6), msg, f["ip"], "example.com", f["pad"][:24], "www.example.com")
            else:
                line = head + msg
        elif k < 0.90:
            t24 = "%s %s %02d %02d:%02d:%02d.%06d 2026" % (_DOW[rng.integers(0, 7)], _MON[rng.integers(0, 12)], rng.integers(1, 29), rng.integers(0, 24), rng.integers(0, 60),
                                                          rng.integers(0, 60), rng.integers(0, 10**6))
            msg = _AMSG[rng.integers(0, len(_AMSG))] + (" " + f["pad"] if rng.random() < 0.3 else "")
            r = rng.random()
            if r < 0.35:
                line = "[%s] [core:error] [pid %d:tid %d] [client %s:%d] AH%05d: %s" % (t24, rng.integers(100, 65000), rng.integers(10**8, 10**9), f["ip"], rng.integers(1024, 65535),
                                                                                     rng.integers(1, 99999), msg)
            elif r < 0.55:
                line = "[%s] [mpm_event:notice] [pid %d:tid %d] AH%05d: %s" % (t24, rng.integers(100, 65000), rng.integers(10**8, 10**9), rng.integers(1, 99999), msg)
            elif r < 0.8:
                line = "[%s] [error] [client %s] %s" % (t24[:19] + " 2026", f["ip"], msg)
            else:
                line = "[%s] [notice] %s" % (t24[:19] + " 2026", msg)
        else:
            r = rng.random()
            if r < 0.3:
                line = "%s - - [%s] GET /%s HTTP/1.1 200 12" % (f["ip"], ts, f["pad"][:20])  # request not quoted
            elif r < 0.6:
                line = "2026-10-17T13:55:36Z service=%s level=info msg=\"%s\"" % (f["pad"][:12], _MSG[rng.integers(0, len(_MSG))].replace('"', "'"))
            elif r < 0.8:
                line = "[%s [error] [client %s] truncated" % (ts, f["ip"])
            else:
                line = _pad_token(rng, alnum + " ", int(rng.integers(0, 80)))
        out.append(line)
    return out


def lines_to_text(lines) -> np.ndarray:
    """'\\n'-terminated UTF-16 text (uint16) of a list of Python strings (lone surrogates pass through)."""
    s = "\n".join(lines) + "\n" if lines else ""
    return np.frombuffer(s.encode("utf-16-le", "surrogatepass"), dtype=np.uint16).copy()


def weblog_corpus(n_lines: int, seed: int = 0x5EED0003) -> np.ndarray:
    return lines_to_text(weblog_lines(n_lines, seed))


# ---------------------------------------------------------------------------------------------------------------
# config #4: 200 syslog-style extractions sharing `<%num>$ts(%phrase) $host(%phrase) ` then `app_k[$pid(%num)]: ` and
# message template k with 2..6 extractors. The combined DFA outgrows shared memory (L2-resident table path).
_W4 = ["accepted", "connection", "from", "port", "session", "opened", "closed", "for", "user", "failed", "password", "invalid",
       "request", "timeout", "after", "retry", "queue", "depth", "latency", "bytes", "sent", "received", "cache", "miss", "hit",
       "worker", "started", "stopped", "signal", "code", "exit", "status", "remote", "local", "peer", "reset", "by", "to"]
_P4 = ["%num", "%word", "%phrase", "%ip"]


def syslog200_spec(n_ext: int = 200, seed: int = 0x5EED0004):
    """[(app, [(literal words, extractor name, pattern) ...], trailing words)] for extraction k = 0..n_ext-1."""
    rng = np.random.default_rng(seed)
    spec = []
    for k in range(n_ext):
        parts = []
        for j in range(int(rng.integers(2, 7))):
            words = [_W4[i] for i in rng.integers(0, len(_W4), size=int(rng.integers(1, 3)))]
            parts.append((words, "f%d" % j, _P4[int(rng.integers(0, len(_P4)))]))
        tail = [_W4[i] for i in rng.integers(0, len(_W4), size=int(rng.integers(0, 2)))]
        spec.append(("app_%d" % k, parts, tail))
    return spec


def syslog200_definition(n_ext: int = 200, seed: int = 0x5EED0004) -> str:
    d = ["pattern %num \\d+", "pattern %word \\w+", "pattern %phrase \\S+", "pattern %ip [0-9a-fA-F\\.:]+",
         "template @head <%num>$ts(%phrase) $host(%phrase)"]
    for app, parts, tail in syslog200_spec(n_ext, seed):
        body = " ".join("%s $%s(%s)" % (" ".join(w), name, pat) for w, name, pat in parts)
        if tail:
            body += " " + " ".join(tail)
        d += ["extract %s {" % app.replace("app_", "App"), "  template @head %s[$pid(%%num)]: %s" % (app, body), "}"]
    return "\n".join(d) + "\n"


SYSLOG200_DEF = syslog200_definition()


def syslog200_lines(n_lines: int, seed: int = 0x5EED0004, n_ext: int = 200, special=None):
    """Config #4 lines: uniform over the templates, 5 % non-matching (unknown app, missing field, trailing junk)."""
    spec = syslog200_spec(n_ext)
    rng = np.random.default_rng(seed + 1)
    alnum = "abcdefghijklmnopqrstuvwxyz0123456789"
    out = []
    for i in range(n_lines):
        app, parts, tail = spec[int(rng.integers(0, n_ext))]
        f = {"pad": "", "ip": "%d.%d.%d.%d" % tuple(rng.integers(1, 255, size=4)), "user": _pad_token(rng, alnum, int(rng.integers(3, 10)))}
        if special is not None:
            special(rng, i, f)
        vals = []
        for w, name, pat in parts:
            v = {"%num": str(rng.integers(0, 10**6)), "%word": f["user"], "%phrase": "/" + f["user"] + "/" + _pad_token(rng, alnum, int(rng.integers(2, 9))) + f["pad"],
                 "%ip": f["ip"]}[pat]
            vals.append(" ".join(w) + " " + v)
        line = "<%d>2026-10-17T%02d:%02d:%02d.%03dZ host-%s %s[%d]: %s" % (rng.integers(0, 192), rng.integers(0, 24), rng.integers(0, 60), rng.integers(0, 60), rng.integers(0, 1000),
                                                                         _pad_token(rng, alnum, 5), app, rng.integers(1, 65536), " ".join(vals))
        if tail:
            line += " " + " ".join(tail)
        r = rng.random()
        if r < 0.02:
            line = line.replace(app + "[", "daemon_%d[" % rng.integers(0, 50), 1)
        elif r < 0.035:
            line = line[: max(10, len(line) - int(rng.integers(3, 20)))]
        elif r < 0.05:
            line += " trailing junk"
        out.append(line)
    return out


def syslog200_corpus(n_lines: int, seed: int = 0x5EED0004) -> np.ndarray:
    return lines_to_text(syslog200_lines(n_lines, seed))


# ---------------------------------------------------------------------------------------------------------------
# config #5: the config #3 family with non-ASCII UTF-16 inside the free fields, the DFA-vs-JDK divergence characters
# and 10 KB outlier lines. Content of line i depends only on (seed, i): any sharding sees the same lines.
_NONASCII = ["".join(map(chr, w)) for w in ([0xE9, 0xE8, 0xFC, 0xF1], [0x391, 0x3B2, 0x3B3, 0x3B4], [0x416, 0x438, 0x432, 0x43E],
             [0x4E2D, 0x6587, 0x65E5, 0x672C], [0xFF21, 0xFF22])]
_SUPPL = [chr(0x1F600), chr(0x10400) + chr(0x10401), chr(0x2F800)]
_DIVERGE = [chr(0x0B), chr(0x08), chr(0x85), chr(0x2028), chr(0x2029), chr(0xD83D), chr(0xDE00), chr(0x0D)]


def _special5(rng, i, f):
    r = rng.random()
    if r < 0.0001:
        f["pad"] = _pad_token(rng, "abcdefghij0123456789", 5000)      # 10 KB outlier
    elif r < 0.0011:
        f["pad"] += _DIVERGE[int(rng.integers(0, len(_DIVERGE)))] + "x"
    elif r < 0.0111:
        f["pad"] += _SUPPL[int(rng.integers(0, len(_SUPPL)))]
    elif r < 0.10:
        f["pad"] += _NONASCII[int(rng.integers(0, len(_NONASCII)))]
        if rng.random() < 0.3:
            f["user"] = _NONASCII[int(rng.integers(0, len(_NONASCII)))]


def utf16_mix_lines(n_lines: int, seed: int = 0x5EED0005, first_line: int = 0, block: int = 4096):
    """Config #5 lines [first_line, first_line + n_lines): generated in blocks of `block` lines, each block from its own
    counter-based stream (seed, block index), so that every shard of the corpus is reproducible on its own."""
    out = []
    b0, b1 = first_line // block, (first_line + n_lines + block - 1) // block
    for b in range(b0, b1):
        lines = weblog_lines(block, seed=seed * 1000003 + b, special=_special5)
        lo = max(first_line - b * block, 0)
        hi = min(first_line + n_lines - b * block, block)
        out += lines[lo:hi]
    return out


def utf16_mix_corpus(n_lines: int, seed: int = 0x5EED0005, first_line: int = 0) -> np.ndarray:
    return lines_to_text(utf16_mix_lines(n_lines, seed, first_line))


CONFIGS.update({
    "weblog": (WEBLOG_DEF, weblog_corpus),
    "syslog200": (SYSLOG200_DEF, syslog200_corpus),
    "utf16mix": (WEBLOG_DEF, utf16_mix_corpus),
})
