/* ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Plain-C restatement of gorp's per-line hot path, one `extract` per line exactly like a
 * Java caller's loop, for (a) parity checks at millions of lines and (b) bench.py's CPU
 * baseline ("C restatement of the reference CPU path; JVM unavailable").
 *
 *   step 1  PolyMatcher.match      autom/PolyMatcher.java:123-133
 *           Automata.step/accept   autom/Automata.java:133-139
 *             p = _transitions[p*_stride + _alphabet[c]]; -1 => NO_MATCH; accept[p][0]
 *   step 2  Gorp.extract dispatch  Gorp.java:159-177  (first index wins; null => throw)
 *   step 3  JDKRegexpCookedExtraction.match/_constructMatch
 *                                  jdkre/JDKRegexpCookedExtraction.java:36-59
 *           = java.util.regex backtracking, anchored at both ends, group(1..n) spans.
 *
 * The tables and the backtracking program are produced by the Python oracle
 * (oracle/brics.py, oracle/jdkre.py); nothing here is shared with gorp_b200/.
 * No SIMD, no batching: scalar code, static contiguous partition of lines over threads.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { OP_CHAR = 0, OP_SET = 1, OP_SPLIT = 2, OP_JMP = 3, OP_SAVE = 4, OP_MATCH = 5 };

typedef struct { int32_t pc; int32_t pos; } Ent;

typedef struct {
    const int32_t *alphabet, *trans, *accept_first;
    int stride;
    const int32_t *ops, *op_off, *ngroups, *set_hdr, *set_iv;
    const uint8_t *latin1;   /* [nsets][256] membership of U+0000..U+00FF (BitClass-like fast path) */
    const uint16_t *text;
    const int64_t *starts, *ends;
    int32_t *ext, *spans;
    int span_stride;
    int64_t lo, hi;
    int rc;
} Job;

static int set_contains(const Job *J, int sid, uint32_t cp) {
    if (cp < 256) return J->latin1[(size_t)sid * 256 + cp];
    const int32_t *h = J->set_hdr + 3 * sid;
    const int32_t *iv = J->set_iv + 2 * (size_t)h[1];
    int lo = 0, hi = h[2] - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        if ((uint32_t)iv[2 * mid + 1] < cp) lo = mid + 1;
        else if ((uint32_t)iv[2 * mid] > cp) hi = mid - 1;
        else return 1;
    }
    return 0;
}

/* Matcher.matches(): first full-consumption path in backtracking preference order. */
static int bt_match(const Job *J, const int32_t *ops, const uint16_t *u, int64_t n,
                    int32_t *caps, Ent **stk, size_t *cap) {
    size_t sp = 0;
    int32_t pc = 0;
    int64_t i = 0;
    for (;;) {
        const int32_t *o = ops + 3 * (size_t)pc;
        switch (o[0]) {
        case OP_CHAR:
            if (i < n && u[i] == (uint16_t)o[1]) { i++; pc++; continue; }
            break;
        case OP_SET: {
            if (i >= n) break;
            const int32_t *h = J->set_hdr + 3 * o[1];
            uint32_t cp = u[i];
            int w = 1;
            if (h[0] && cp >= 0xD800 && cp <= 0xDBFF && i + 1 < n) {  /* Character.codePointAt */
                uint32_t d = u[i + 1];
                if (d >= 0xDC00 && d <= 0xDFFF) { cp = 0x10000 + ((cp - 0xD800) << 10) + (d - 0xDC00); w = 2; }
            }
            if (set_contains(J, o[1], cp)) { i += w; pc++; continue; }
            break;
        }
        case OP_SPLIT:
            if (sp + 2 > *cap) {
                *cap *= 2;
                *stk = (Ent *)realloc(*stk, *cap * sizeof(Ent));
                if (!*stk) return -1;
            }
            (*stk)[sp].pc = o[2]; (*stk)[sp].pos = (int32_t)i; sp++;
            pc = o[1];
            continue;
        case OP_JMP:
            pc = o[1];
            continue;
        case OP_SAVE:
            if (sp + 2 > *cap) {
                *cap *= 2;
                *stk = (Ent *)realloc(*stk, *cap * sizeof(Ent));
                if (!*stk) return -1;
            }
            (*stk)[sp].pc = -o[1] - 1; (*stk)[sp].pos = caps[o[1]]; sp++;
            caps[o[1]] = (int32_t)i;
            pc++;
            continue;
        case OP_MATCH:
            if (i == n) return 1;
            break;
        default:
            return -1;
        }
        /* fail: backtrack */
        for (;;) {
            if (sp == 0) return 0;
            Ent e = (*stk)[--sp];
            if (e.pc < 0) { caps[-e.pc - 1] = e.pos; continue; }
            pc = e.pc; i = e.pos;
            break;
        }
    }
}

/* JDKRegexpCookedExtraction._constructMatch (jdkre/JDKRegexpCookedExtraction.java:51-59) materialises every group as
 * a String (values[i] = m.group(i+1): an allocation + a copy of the group's units). The baseline pays for the copy too:
 * each group's units are copied into a per-thread arena (a ring, so the working set stays cache-sized like a young-generation
 * allocation buffer). The copies do not change the reported spans. */
#define ARENA_UNITS (1u << 18)

static void *worker(void *arg) {
    Job *J = (Job *)arg;
    size_t cap = 1024;
    Ent *stk = (Ent *)malloc(cap * sizeof(Ent));
    uint16_t *arena = (uint16_t *)malloc(ARENA_UNITS * sizeof(uint16_t));
    size_t arena_pos = 0;
    volatile uint16_t sink = 0;
    int32_t caps[256];
    for (int64_t l = J->lo; l < J->hi; l++) {
        const uint16_t *u = J->text + J->starts[l];
        int64_t n = J->ends[l] - J->starts[l];
        int32_t *out = J->spans + (size_t)l * J->span_stride;
        /* PolyMatcher.match */
        int32_t p = 0;
        int64_t i = 0;
        for (; i < n; i++) {
            p = J->trans[(size_t)p * J->stride + J->alphabet[u[i]]];
            if (p == -1) break;
        }
        int32_t e = (p == -1) ? -1 : J->accept_first[p];
        if (e < 0) { J->ext[l] = -1; continue; }
        /* CookedExtraction.match */
        int g = J->ngroups[e];
        if (2 * g + 2 > 256) { J->rc = -2; break; }
        for (int k = 0; k < 2 * g + 2; k++) caps[k] = -1;
        int r = bt_match(J, J->ops + 3 * (size_t)J->op_off[e], u, n, caps, &stk, &cap);
        if (r < 0) { J->rc = -1; break; }
        if (r == 0) { J->ext[l] = -2 - e; continue; }
        J->ext[l] = e;
        for (int k = 0; k < 2 * g && k < J->span_stride; k++) out[k] = caps[2 + k];  /* group(1..n) */
        for (int k = 0; k < g; k++) {                                                 /* values[k] = m.group(k + 1) */
            int32_t s0 = caps[2 + 2 * k], s1 = caps[3 + 2 * k];
            if (s0 < 0 || s1 <= s0) continue;
            size_t len = (size_t)(s1 - s0);
            if (len > ARENA_UNITS) len = ARENA_UNITS;
            if (arena_pos + len > ARENA_UNITS) arena_pos = 0;
            memcpy(arena + arena_pos, u + s0, len * sizeof(uint16_t));
            sink ^= arena[arena_pos];
            arena_pos += len;
        }
    }
    (void)sink;
    free(arena);
    free(stk);
    return NULL;
}

int gorp_oracle_run(const int32_t *alphabet, const int32_t *trans, int stride, const int32_t *accept_first,
                    const int32_t *ops, const int32_t *op_off, const int32_t *ngroups,
                    const int32_t *set_hdr, const int32_t *set_iv,
                    const uint16_t *text, const int64_t *starts, const int64_t *ends, int64_t n,
                    int32_t *ext, int32_t *spans, int span_stride, int threads, int n_ext, int n_sets) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    for (int e = 0; e < n_ext; e++)
        if (ops[3 * (size_t)(op_off[e + 1] - 1)] != OP_MATCH) return -3;
    int nsets = n_sets;
    uint8_t *latin1 = (uint8_t *)calloc((size_t)(nsets > 0 ? nsets : 1) * 256, 1);
    for (int s = 0; s < nsets; s++) {
        const int32_t *h = set_hdr + 3 * s;
        const int32_t *iv = set_iv + 2 * (size_t)h[1];
        for (int k = 0; k < h[2]; k++)
            for (int64_t c = iv[2 * k]; c <= iv[2 * k + 1] && c < 256; c++) latin1[(size_t)s * 256 + c] = 1;
    }
    Job *jobs = (Job *)calloc(threads, sizeof(Job));
    pthread_t *tids = (pthread_t *)calloc(threads, sizeof(pthread_t));
    int rc = 0;
    for (int t = 0; t < threads; t++) {
        Job *J = &jobs[t];
        J->alphabet = alphabet; J->trans = trans; J->accept_first = accept_first; J->stride = stride;
        J->ops = ops; J->op_off = op_off; J->ngroups = ngroups; J->set_hdr = set_hdr; J->set_iv = set_iv;
        J->latin1 = latin1; J->text = text; J->starts = starts; J->ends = ends;
        J->ext = ext; J->spans = spans; J->span_stride = span_stride;
        J->lo = n * t / threads; J->hi = n * (t + 1) / threads;
        if (threads == 1) worker(J);
        else pthread_create(&tids[t], NULL, worker, J);
    }
    for (int t = 0; t < threads; t++) {
        if (threads > 1) pthread_join(tids[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    free(jobs); free(tids); free(latin1);
    return rc;
}
