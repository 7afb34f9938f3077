// BENCH / TEST INFRASTRUCTURE (not part of libgorpcuda.so): synthetic log corpora whose line i is a pure function of
// (seed, i) — SURVEY.md §8(d) config #5: "content of line i = f(seed, i) via a counter-based RNG so shards are identical for
// any GPU count; generated on-device". One source for both sides: this header is compiled by nvcc into the device kernels
// and the host loops of libgorpgen.so (csrc/tools/corpusgen.cu), so the text a GPU shard holds and the text the CPU oracle
// checks are the same by construction (integer arithmetic only: no libm, no floating point).
//
// A corpus is described by a small program (built by gorp_b200/corpusgen.py from the line grammars of gorp_b200/corpus.py):
//   kinds   [n_kinds][2]   (cumulative probability as u32, first op of the kind's program)
//   ops     [n_ops][4]     (opcode, a, b, c)
//   choice  [..][3]        (cumulative probability as u32, string offset, string length)
//   strings u16[]          literal pool (UTF-16 units), alphabets
//   qtable  [256]          quantiles of the per-line "pad" length (heavy-tailed field lengths without floating point)
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define CG_HD __host__ __device__ __forceinline__
#else
#define CG_HD static inline
#endif

enum CgOp {
    CG_END = 0,
    CG_LIT = 1,      // a = string offset, b = length
    CG_NUM = 2,      // decimal of a uniform integer in [a, b]
    CG_NUMPAD = 3,   // same, zero-padded to c digits
    CG_TOKEN = 4,    // a = alphabet (offset), b = alphabet length | (min length << 16), c = max length
    CG_CHOICE = 5,   // a = first entry of `choice`, b = entries
    CG_SKIPIF = 6,   // with probability a / 2^32 skip the next b ops
    CG_PAD = 7,      // the line's pad field; a = part (0 all, 1 first half, 2 second half), b = cap on the length (0 = none),
                     // c = 1: followed by the line's special suffix (config #5)
    CG_IP = 8,       // d.d.d.d (90 %) or 2001:db8::x:y
    CG_USER = 9,     // "-" (80 %) or a 6-letter token; config #5: sometimes a non-ASCII word
    CG_HEX = 10,     // lowercase hex of a uniform integer in [a, b]
};

struct CgProgram {
    const uint32_t* kinds;    // [n_kinds * 2]
    const int32_t* ops;       // [n_ops * 4]
    const uint32_t* choice;   // [n * 3]
    const uint16_t* strings;
    const uint16_t* qtable;   // [256]
    uint32_t n_kinds;
    // config #5 specials (all 0 = none): thresholds on one u32 draw per line, ascending
    uint32_t p_outlier, p_diverge, p_suppl, p_nonascii;   // cumulative
    uint32_t outlier_len;
    uint32_t diverge_off, diverge_n;   // `choice` entries (strings) of the divergence characters
    uint32_t suppl_off, suppl_n;
    uint32_t nonascii_off, nonascii_n;
    uint32_t alnum_off, alnum_n;       // alphabet of pad / user tokens
    uint32_t outlier_alpha_off, outlier_alpha_n;
};

struct CgRng {
    uint64_t s;
};
CG_HD uint64_t cg_mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
CG_HD CgRng cg_seed(uint64_t seed, uint64_t line) {
    CgRng r;
    r.s = cg_mix(seed * 0x9E3779B97F4A7C15ull + line * 0xD1B54A32D192ED03ull + 0x632BE59BD9B4E019ull);
    return r;
}
CG_HD uint32_t cg_next(CgRng& r) {
    r.s += 0x9E3779B97F4A7C15ull;
    return static_cast<uint32_t>(cg_mix(r.s) >> 32);
}
CG_HD uint32_t cg_range(CgRng& r, uint32_t lo, uint32_t hi) {  // uniform in [lo, hi]
    return lo + cg_next(r) % (hi - lo + 1u);
}

struct CgOut {
    uint16_t* out;   // null: only count
    int64_t pos;
};
CG_HD void cg_put(CgOut& o, uint32_t unit) {
    if (o.out) o.out[o.pos] = static_cast<uint16_t>(unit);
    ++o.pos;
}
CG_HD void cg_puts(CgOut& o, const uint16_t* s, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) cg_put(o, s[i]);
}
CG_HD void cg_decimal(CgOut& o, uint32_t v, uint32_t width) {
    uint16_t buf[12];
    uint32_t n = 0;
    do {
        buf[n++] = static_cast<uint16_t>('0' + v % 10u);
        v /= 10u;
    } while (v);
    while (n < width) buf[n++] = '0';
    while (n) cg_put(o, buf[--n]);
}
CG_HD void cg_hex(CgOut& o, uint32_t v) {
    uint16_t buf[8];
    uint32_t n = 0;
    do {
        const uint32_t d = v & 15u;
        buf[n++] = static_cast<uint16_t>(d < 10 ? '0' + d : 'a' + d - 10);
        v >>= 4;
    } while (v);
    while (n) cg_put(o, buf[--n]);
}
CG_HD void cg_token(CgOut& o, CgRng& r, const uint16_t* alpha, uint32_t n_alpha, uint32_t len) {
    for (uint32_t i = 0; i < len; ++i) cg_put(o, alpha[cg_next(r) % n_alpha]);
}
CG_HD void cg_choice_string(CgOut& o, CgRng& r, const CgProgram& P, uint32_t first, uint32_t n) {
    const uint32_t x = cg_next(r);
    uint32_t i = 0;
    while (i + 1 < n && x >= P.choice[(first + i) * 3]) ++i;
    cg_puts(o, P.strings + P.choice[(first + i) * 3 + 1], P.choice[(first + i) * 3 + 2]);
}

// Writes line `line` (without its '\n') at out (or only counts when out == null). Returns the length in units.
CG_HD int64_t cg_line(const CgProgram& P, uint64_t seed, uint64_t line, uint16_t* out) {
    CgRng r = cg_seed(seed, line);
    CgOut o{out, 0};
    // per-line draws, always in the same order
    const uint32_t kind_x = cg_next(r), special_x = cg_next(r), pad_q = cg_next(r), aux = cg_next(r);
    uint32_t pad_len = P.qtable[pad_q & 255u];
    uint32_t special = 0;  // 1 outlier, 2 diverge, 3 suppl, 4 nonascii
    if (P.p_nonascii) {
        if (special_x < P.p_outlier) special = 1;
        else if (special_x < P.p_diverge) special = 2;
        else if (special_x < P.p_suppl) special = 3;
        else if (special_x < P.p_nonascii) special = 4;
    }
    if (special == 1) pad_len = P.outlier_len;
    const uint64_t pad_seed = cg_mix(r.s ^ 0xA5A5A5A5A5A5A5A5ull);  // the pad field's characters: its own stream (it may be split)
    uint32_t k = 0, hi = P.n_kinds - 1;  // first kind whose cumulative probability exceeds the draw
    while (k < hi) {
        const uint32_t mid = (k + hi) >> 1;
        if (kind_x >= P.kinds[2 * mid]) k = mid + 1;
        else hi = mid;
    }
    const int32_t* op = P.ops + 4 * P.kinds[2 * k + 1];
    for (;; op += 4) {
        switch (op[0]) {
            case CG_END: return o.pos;
            case CG_LIT: cg_puts(o, P.strings + op[1], static_cast<uint32_t>(op[2])); break;
            case CG_NUM: cg_decimal(o, cg_range(r, static_cast<uint32_t>(op[1]), static_cast<uint32_t>(op[2])), 0); break;
            case CG_NUMPAD: cg_decimal(o, cg_range(r, static_cast<uint32_t>(op[1]), static_cast<uint32_t>(op[2])), static_cast<uint32_t>(op[3])); break;
            case CG_TOKEN: {
                const uint32_t n_alpha = static_cast<uint32_t>(op[2]) & 0xFFFFu, mn = static_cast<uint32_t>(op[2]) >> 16;
                cg_token(o, r, P.strings + op[1], n_alpha, cg_range(r, mn, static_cast<uint32_t>(op[3])));
                break;
            }
            case CG_CHOICE: cg_choice_string(o, r, P, static_cast<uint32_t>(op[1]), static_cast<uint32_t>(op[2])); break;
            case CG_SKIPIF:
                if (cg_next(r) < static_cast<uint32_t>(op[1])) op += 4 * op[2];
                break;
            case CG_PAD: {
                uint32_t lo = 0, hi = pad_len;
                if (op[1] == 1) hi = special == 1 ? pad_len : pad_len / 2;   // an outlier line keeps its 10 KB field in one piece
                if (op[1] == 2) lo = special == 1 ? pad_len : pad_len / 2;
                if (op[2] && hi - lo > static_cast<uint32_t>(op[2])) hi = lo + static_cast<uint32_t>(op[2]);
                const uint16_t* alpha = P.strings + (special == 1 ? P.outlier_alpha_off : P.alnum_off);
                const uint32_t n_alpha = special == 1 ? P.outlier_alpha_n : P.alnum_n;
                for (uint32_t i = lo; i < hi; ++i) cg_put(o, alpha[static_cast<uint32_t>(cg_mix(pad_seed + i) >> 33) % n_alpha]);
                if (op[3] == 1 && special >= 2) {
                    const uint32_t first = special == 2 ? P.diverge_off : (special == 3 ? P.suppl_off : P.nonascii_off);
                    const uint32_t n = special == 2 ? P.diverge_n : (special == 3 ? P.suppl_n : P.nonascii_n);
                    const uint32_t i = (aux >> 8) % n;
                    cg_puts(o, P.strings + P.choice[(first + i) * 3 + 1], P.choice[(first + i) * 3 + 2]);
                    if (special == 2) cg_put(o, 'x');
                }
                break;
            }
            case CG_IP:
                if (cg_next(r) % 10u) {
                    for (int i = 0; i < 4; ++i) {
                        if (i) cg_put(o, '.');
                        cg_decimal(o, cg_range(r, 1, 254), 0);
                    }
                } else {
                    const uint16_t pre[10] = {'2', '0', '0', '1', ':', 'd', 'b', '8', ':', ':'};
                    cg_puts(o, pre, 10);
                    cg_hex(o, cg_range(r, 1, 65534));
                    cg_put(o, ':');
                    cg_hex(o, cg_range(r, 1, 65534));
                }
                break;
            case CG_USER:
                if (special == 4 && (aux & 0xFFu) < 77u) {  // 30 % of the non-ASCII lines: a non-ASCII user name
                    const uint32_t i = (aux >> 16) % P.nonascii_n;
                    cg_puts(o, P.strings + P.choice[(P.nonascii_off + i) * 3 + 1], P.choice[(P.nonascii_off + i) * 3 + 2]);
                } else if (cg_next(r) % 5u) {
                    cg_put(o, '-');
                } else {
                    cg_token(o, r, P.strings + P.alnum_off, P.alnum_n, 6);
                }
                break;
            case CG_HEX: cg_hex(o, cg_range(r, static_cast<uint32_t>(op[1]), static_cast<uint32_t>(op[2]))); break;
            default: return o.pos;
        }
    }
}
