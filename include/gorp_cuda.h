/* libgorpcuda — C ABI of the B200 batch extraction engine for gorp definitions.
 *
 * Drop-in boundary. Each entry point names the reference interface it replaces; paths are relative
 * to gorp-core/src/main/java/com/salesforce/gorp/ of salesforce/gorp.
 *
 *   reference (per line, JVM)                                  this library (per batch, GPU)
 *   --------------------------------------------------------   ------------------------------------------------
 *   DefinitionReader.reader(..).read()   DefinitionReader.java:51-84
 *   Gorp.construct(defs, cooker)         Gorp.java:50-92        gorp_compile_definition   (JVM-less hosts)
 *   Automata fields + regexp sources     autom/Automata.java:23-26,
 *                                        model/CookedExtraction.java:46   -> DfaExport blob -> gorp_engine_create
 *   PolyMatcher.match(CharSequence)      autom/PolyMatcher.java:123-133  \
 *   Gorp.extract(String)                 Gorp.java:145-186                > gorp_extract_lines / gorp_extract_text
 *   CookedExtraction.match(String)       jdkre/JDKRegexpCookedExtraction.java:36-59 /
 *   new ExtractionResult(...)            model/CookedExtraction.java:54-57   (Java materialises from gorp_result)
 *
 * Conventions: plain pointers and sizes, no exceptions cross the boundary; 0 = OK, < 0 = gorp_status; the message of
 * the last failure on the calling thread is gorp_last_error(). Text is UTF-16 code units in host byte order (Java
 * `char`). There is NO CPU fallback: every gorp_extract_* call fails with GORP_E_CUDA when no usable device exists.
 */
#ifndef GORP_CUDA_H
#define GORP_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GORP_ABI_VERSION 2

typedef enum gorp_status {
    GORP_OK = 0,
    GORP_E_ARG = -1,         /* null / out-of-range argument */
    GORP_E_DEFINITION = -2,  /* DefinitionParseException of the reference (DefinitionParseException.java:17) */
    GORP_E_UNSUPPORTED = -3, /* definition is legal for the reference but outside the GPU subset */
    GORP_E_BLOB = -4,        /* malformed / corrupt DfaExport blob */
    GORP_E_CUDA = -5,        /* CUDA error or no device */
    GORP_E_OOM = -6,
    GORP_E_INTERNAL = -7
} gorp_status;

/* Per-line outcome codes in gorp_result.ext_id (Gorp.java:159-177):
 *   >= 0        index of the matched extraction (declaration order), spans present
 *   GORP_MISS   extract() would return null
 *   <= -2       extract() would throw ExtractionException for extraction index (-2 - ext_id):
 *               the combined DFA accepted but the java.util.regex dialect of that extraction did not */
#define GORP_MISS (-1)
#define GORP_CAPTURE_FAIL(e) (-2 - (e))

typedef struct gorp_engine gorp_engine;

typedef struct gorp_result {
    int64_t n_lines;
    int32_t n_extractions;
    int32_t span_stride;      /* int32 entries per result row of `spans` = 2 * (groups of the widest extraction) */
    const int32_t* ext_id;    /* [n_lines] */
    const int64_t* line_off;  /* [n_lines+1] start of each line in the caller's text; line i spans
                                 [line_off[i], line_off[i+1] - sep) with sep = 1 for gorp_extract_text ('\n'
                                 separated; line_off[n_lines] = n_units (+1 when the text does not end in '\n'))
                                 and sep = 0 for gorp_extract_lines (a copy of the caller's offsets) */
    const int32_t* spans;     /* [n_lines * span_stride] one fixed-size row per line: row i starts at i * span_stride
                                 and holds 2 * n_groups(ext_id[i]) valid entries = (start, end) pairs in UTF-16 units
                                 relative to the start of the line, in group order == extractor-name order; (-1,-1)
                                 when a group did not participate; the rest of the row, and the whole row of a MISS
                                 or capture-failed line, is -1. Fixed rows (instead of a CSR) keep the device path
                                 free of a span-offset prefix sum and let Java index rows directly. */
    const int64_t* histogram; /* [n_extractions + 2]: lines per extraction, then MISS, then capture failures */
    void* owner;              /* internal */
} gorp_result;

typedef struct gorp_blob_info {
    uint32_t n_states, n_classes, n_extractions, reserved;
} gorp_blob_info;

typedef struct gorp_extraction_info {  /* pointers stay valid as long as the blob they came from */
    uint32_t n_groups;
    uint32_t n_extractor_names;
    const uint16_t* name;            uint32_t name_len;
    const uint16_t* automaton_regex; uint32_t automaton_regex_len;  /* brics dialect, Gorp.java:68 */
    const uint16_t* jdk_regex;       uint32_t jdk_regex_len;        /* CookedExtraction.getRegexpSource() */
    const char* append_json;         uint32_t append_json_len;      /* raw `append` objects, '\n' separated */
} gorp_extraction_info;

int gorp_abi_version(void);
const char* gorp_last_error(void);
int gorp_device_count(void);

/* --- definition -> blob (host only; what DefinitionReader.read() + Gorp.construct() + DfaExport do in Java) */
int gorp_compile_definition(const char* grp_utf8, size_t len, void** blob, size_t* blob_len);
/* PolyMatcher.create(String...) for raw automaton-dialect patterns (autom/PolyMatcher.java:64-84): a blob whose
 * extractions have no capture groups (the JDK dialect string is left empty => match-only engine). */
int gorp_compile_patterns(const uint16_t* const* patterns, const uint32_t* lens, uint32_t n, void** blob, size_t* blob_len);
void gorp_blob_free(void* blob);

/* --- blob introspection (host only) */
int gorp_blob_get_info(const void* blob, size_t len, gorp_blob_info* out);
int gorp_blob_get_extraction(const void* blob, size_t len, uint32_t index, gorp_extraction_info* out);
int gorp_blob_get_extractor_name(const void* blob, size_t len, uint32_t extraction, uint32_t k,
                                 const uint16_t** name, uint32_t* name_len);
/* Raw tables exactly as Automata holds them (autom/Automata.java:23-26); accept lists as CSR. */
int gorp_blob_get_tables(const void* blob, size_t len, const uint16_t** classmap, const int32_t** transitions,
                         const int32_t** accept_first, const uint32_t** accept_off, const int32_t** accept_list);

/* --- engine: immutable after creation; concurrent gorp_extract_* calls on one engine are allowed */
int gorp_engine_create(const void* blob, size_t len, const int* devices, int n_devices, gorp_engine** out);
void gorp_engine_destroy(gorp_engine* e);

/* Gorp.extractAll(List<String>): `text` is the concatenation of the strings, off[i]..off[i+1] delimit string i. */
int gorp_extract_lines(gorp_engine* e, const uint16_t* text, const int64_t* off, int64_t n_lines, gorp_result* out);
/* Gorp.extractAll(CharBuffer): split on U+000A only; a final line without '\n' counts; '\r' is data. */
int gorp_extract_text(gorp_engine* e, const uint16_t* text, int64_t n_units, gorp_result* out);
/* Same call for text held as ISO-8859-1 bytes, one byte per character — what a JDK 9+ String with the LATIN1 coder holds
 * (java.lang.String compact strings; `String.getBytes(ISO_8859_1)` is a plain copy for such strings), i.e. the usual
 * case for log text. The bytes are widened to UTF-16 on the device, so the host-to-device copy moves half the bytes of
 * gorp_extract_text; results are identical to gorp_extract_text on the zero-extended text (spans and line_off count
 * characters == UTF-16 units). Split on byte 0x0A. */
int gorp_extract_text_latin1(gorp_engine* e, const uint8_t* text, int64_t n_bytes, gorp_result* out);
/* Same call for text held as UTF-8 bytes — log files as they are on disk (the reference reads UTF-8 the same way for its
 * definitions, io/InputLineReader.java:51; a Java caller would otherwise decode the bytes to Strings first). The bytes are
 * decoded to UTF-16 on the device (supplementary characters become surrogate pairs), so ASCII text crosses PCIe at one byte
 * per character. Results are those of gorp_extract_text on `new String(bytes, UTF_8)`: line_off and spans count UTF-16
 * units of the DECODED text. Split on byte 0x0A. The input must be well-formed UTF-8 (Unicode table 3-7): a malformed
 * sequence fails the call with GORP_E_ARG and the byte offset in gorp_last_error() — no replacement characters are
 * invented. Runs on the engine's first device. */
int gorp_extract_text_utf8(gorp_engine* e, const uint8_t* text, int64_t n_bytes, gorp_result* out);
void gorp_result_release(gorp_engine* e, gorp_result* r);

/* PolyMatcher.match for a batch (reference autom/PolyMatcher.java:123-133, reached through Gorp.getMatcher(), Gorp.java:135-137):
 * for every string ALL regex / extraction indexes whose automaton accepts the whole string, ascending (Automata.accept,
 * autom/Automata.java:137-139) — not only the first one the extract path uses. CSR: the indexes of string i are
 * accept[accept_off[i] .. accept_off[i+1]). Works on any engine (definitions and gorp_compile_patterns blobs). */
typedef struct gorp_match_result {
    int64_t n_lines;
    const int64_t* accept_off; /* [n_lines + 1] */
    const int32_t* accept;     /* [accept_off[n_lines]] */
    void* owner;
} gorp_match_result;
int gorp_match_all_lines(gorp_engine* e, const uint16_t* text, const int64_t* off, int64_t n_lines, gorp_match_result* out);
void gorp_match_result_release(gorp_engine* e, gorp_match_result* r);

/* --- device-resident variants: `d_text` (and `d_off`) already live in the HBM of the engine's device `dev_index`
 * (index into the `devices` array given at creation) and must be 16-byte aligned; results stay in device memory
 * owned by the engine until the next call on the same (engine, dev_index). `stream` is a cudaStream_t (NULL = the
 * CUDA default stream, as everywhere in CUDA). The call enqueues the kernels on that stream; the text form waits once for the newline
 * count (it sizes the per-line arrays). Used by bench.py for the HBM-resident `value`. */
#define GORP_FLAG_SYNC 1          /* cudaStreamSynchronize before returning */
#define GORP_FLAG_TIME_KERNELS 2  /* bracket every kernel with CUDA events on `stream` (see gorp_kernel_times) */

typedef struct gorp_device_result {
    int64_t n_lines;
    int32_t span_stride;
    int32_t reserved;
    const int32_t* d_ext_id;
    const int64_t* d_line_off;
    const int32_t* d_spans;
    const int64_t* d_histogram;
    const int64_t* d_n_lines; /* device scalar */
} gorp_device_result;

int gorp_extract_text_device(gorp_engine* e, int dev_index, const uint16_t* d_text, int64_t n_units, void* stream,
                             int flags, gorp_device_result* out);
int gorp_extract_lines_device(gorp_engine* e, int dev_index, const uint16_t* d_text, const int64_t* d_off,
                              int64_t n_lines, void* stream, int flags, gorp_device_result* out);

/* Device time per pipeline stage, summed over the calls made with GORP_FLAG_TIME_KERNELS on (e, dev_index) since the
 * last reset: names[i] / total_ms[i] for i < *n (at most `cap`), `calls` = timed calls, `launches` = kernels launched
 * by this context (all calls). Waits for the last timed call to finish. */
int gorp_kernel_times(gorp_engine* e, int dev_index, const char** names, double* total_ms, int cap, int* n,
                      int64_t* calls, int64_t* launches, int reset);

#ifdef __cplusplus
}
#endif
#endif /* GORP_CUDA_H */
