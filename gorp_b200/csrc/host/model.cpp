#include "model.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace gorp {

CompiledDefinition compile_definition(const ustring& grp_text) {
    CompiledDefinition out;
    std::vector<ExtractionStrings> xs = read_definition(grp_text);
    // Gorp.construct order: every extraction's java.util.regex source is compiled first (Gorp.java:58-79) ...
    for (auto& x : xs) {
        CompiledExtraction ce;
        try {
            CaptureProgram p = compile_jdk_regex(x.jdk_regex);
            ce.n_groups = static_cast<uint32_t>(p.n_groups);
        } catch (const DefinitionParseError& e) {
            throw DefinitionParseError(std::string("Internal problem: invalid regular expression segment, problem: ") + e.what());
        }
        if (ce.n_groups != x.extractor_names.size())
            throw UnsupportedError(strfmt("extraction '%s': %u capturing groups for %zu extractors", utf16_to_utf8(x.name).c_str(),
                                          ce.n_groups, x.extractor_names.size()));
        ce.strings = std::move(x);
        out.extractions.push_back(std::move(ce));
    }
    // ... then the poly-matcher is built (Gorp.java:81-90)
    std::vector<ustring> pats;
    for (auto& e : out.extractions) pats.push_back(e.strings.automaton_regex);
    out.dfa = compile_patterns(pats);
    return out;
}

DfaTables compile_patterns(const std::vector<ustring>& automaton_regexes) {
    std::vector<MinDfa> dfas;
    for (auto& p : automaton_regexes) {
        try {
            dfas.push_back(brics_min_dfa(p));
        } catch (const std::invalid_argument& e) {
            throw DefinitionParseError(std::string("Internal error: problem with PolyMatcher construction: Invalid regexp, ") +
                                       e.what() + ", source: " + utf16_to_utf8(p));
        }
    }
    return build_product(dfas);
}

// ------------------------------------------------------------------ blob
namespace {
constexpr char kMagic[8] = {'G', 'O', 'R', 'P', 'D', 'F', 'A', '1'};
constexpr size_t kHeader = 8 + 6 * 4 + 8;

uint64_t fnv1a(const uint8_t* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

struct Writer {
    std::vector<uint8_t> b;
    void raw(const void* p, size_t n) {
        const uint8_t* q = static_cast<const uint8_t*>(p);
        b.insert(b.end(), q, q + n);
    }
    void u32(uint32_t v) { raw(&v, 4); }
    void pad() {
        while (b.size() % 4) b.push_back(0);
    }
    void str(const ustring& s) {
        u32(static_cast<uint32_t>(s.size()));
        raw(s.data(), s.size() * 2);
        pad();
    }
};

struct Cursor {
    const uint8_t* p;
    size_t n, i = 0;
    void need(size_t k) const {
        if (k > n - i) throw BlobError("blob truncated");
    }
    uint32_t u32() {
        need(4);
        uint32_t v;
        std::memcpy(&v, p + i, 4);
        i += 4;
        return v;
    }
    template <class T>
    void vec(std::vector<T>& out, size_t count) {
        if (count > (n - i) / sizeof(T)) throw BlobError("blob truncated");
        out.resize(count);
        std::memcpy(out.data(), p + i, count * sizeof(T));
        i += count * sizeof(T);
    }
    void pad() {
        while (i % 4) ++i;
    }
    ustring str() {
        uint32_t len = u32();
        need(static_cast<size_t>(len) * 2);
        ustring s(len, u'\0');
        std::memcpy(s.data(), p + i, static_cast<size_t>(len) * 2);
        i += static_cast<size_t>(len) * 2;
        pad();
        return s;
    }
};
}  // namespace

std::vector<uint8_t> serialize_blob(const CompiledDefinition& d) {
    Writer w;
    w.raw(kMagic, 8);
    w.u32(1);
    w.u32(d.dfa.n_states);
    w.u32(d.dfa.n_classes);
    w.u32(static_cast<uint32_t>(d.extractions.size()));
    w.u32(0);
    w.u32(0);
    uint64_t zero = 0;
    w.raw(&zero, 8);
    w.raw(d.dfa.classmap.data(), 65536 * 2);
    w.raw(d.dfa.trans.data(), d.dfa.trans.size() * 4);
    w.raw(d.dfa.accept_first.data(), d.dfa.accept_first.size() * 4);
    w.raw(d.dfa.accept_off.data(), d.dfa.accept_off.size() * 4);
    w.raw(d.dfa.accept_list.data(), d.dfa.accept_list.size() * 4);
    for (auto& e : d.extractions) {
        w.u32(e.n_groups);
        w.str(e.strings.name);
        w.str(e.strings.automaton_regex);
        w.str(e.strings.jdk_regex);
        w.u32(static_cast<uint32_t>(e.strings.extractor_names.size()));
        for (auto& n : e.strings.extractor_names) w.str(n);
        w.u32(static_cast<uint32_t>(e.strings.append_json.size()));
        w.raw(e.strings.append_json.data(), e.strings.append_json.size());
        w.pad();
    }
    uint64_t sum = fnv1a(w.b.data() + kHeader, w.b.size() - kHeader);
    std::memcpy(w.b.data() + kHeader - 8, &sum, 8);
    return std::move(w.b);
}

CompiledDefinition parse_blob(const void* data, size_t len) {
    if (!data || len < kHeader) throw BlobError("blob too short");
    Cursor c{static_cast<const uint8_t*>(data), len};
    if (std::memcmp(c.p, kMagic, 8) != 0) throw BlobError("bad magic (expected GORPDFA1)");
    c.i = 8;
    if (c.u32() != 1) throw BlobError("unsupported blob version");
    CompiledDefinition d;
    d.dfa.n_states = c.u32();
    d.dfa.n_classes = c.u32();
    uint32_t E = c.u32();
    c.u32();
    c.u32();
    uint64_t sum;
    std::memcpy(&sum, c.p + c.i, 8);
    c.i += 8;
    if (sum != fnv1a(c.p + kHeader, len - kHeader)) throw BlobError("checksum mismatch");
    const size_t S = d.dfa.n_states, C = d.dfa.n_classes;
    if (S == 0 || C == 0 || C > 65536 || S > (1u << 26)) throw BlobError("implausible table dimensions");
    d.dfa.n_regex = E;
    c.vec(d.dfa.classmap, 65536);
    c.vec(d.dfa.trans, S * C);
    c.vec(d.dfa.accept_first, S);
    c.vec(d.dfa.accept_off, S + 1);
    if (d.dfa.accept_off[0] != 0) throw BlobError("bad accept CSR");
    for (size_t s = 0; s < S; ++s)
        if (d.dfa.accept_off[s + 1] < d.dfa.accept_off[s]) throw BlobError("bad accept CSR");
    c.vec(d.dfa.accept_list, d.dfa.accept_off[S]);
    for (uint16_t k : d.dfa.classmap)
        if (k >= C) throw BlobError("classmap entry out of range");
    for (int32_t t : d.dfa.trans)
        if (t < -1 || t >= static_cast<int32_t>(S)) throw BlobError("transition out of range");
    for (int32_t a : d.dfa.accept_first)
        if (a < -1 || a >= static_cast<int32_t>(E)) throw BlobError("accept index out of range");
    for (uint32_t e = 0; e < E; ++e) {
        CompiledExtraction x;
        x.n_groups = c.u32();
        x.strings.name = c.str();
        x.strings.automaton_regex = c.str();
        x.strings.jdk_regex = c.str();
        uint32_t nn = c.u32();
        if (nn > 4096) throw BlobError("implausible extractor count");
        for (uint32_t k = 0; k < nn; ++k) x.strings.extractor_names.push_back(c.str());
        uint32_t jl = c.u32();
        c.need(jl);
        x.strings.append_json.assign(reinterpret_cast<const char*>(c.p + c.i), jl);
        c.i += jl;
        c.pad();
        d.extractions.push_back(std::move(x));
    }
    return d;
}

DeviceModel build_device_model(const CompiledDefinition& d) {
    DeviceModel m;
    m.dfa = compact_tables(d.dfa);
    for (auto& e : d.extractions) {
        CaptureProgram p = compile_jdk_regex(e.strings.jdk_regex);
        if (static_cast<uint32_t>(p.n_groups) != e.n_groups) throw BlobError("group count in blob disagrees with the regex");
        m.n_groups.push_back(e.n_groups);
        m.programs.push_back(std::move(p));
    }
    m.symbols = build_symbol_classes(m.programs);
    size_t max_states = 60000;
    if (const char* f = std::getenv("GORP_TDFA_MAX_STATES")) max_states = static_cast<size_t>(std::max(2, std::atoi(f)));  // tests
    m.pike.resize(m.programs.size());
    for (size_t e = 0; e < m.programs.size(); ++e) {
        try {
            m.tdfas.push_back(build_tdfa(m.programs[e], m.symbols, max_states));
            m.pike_only.push_back(0);
        } catch (const UnsupportedError&) {
            // the determinisation outgrew the limits (states / registers / command lists): the reference accepts any regex
            // Pattern.compile accepts (jdkre/JDKRegexpExtractionCooker.java:20-26), so the extraction falls back to a
            // simulated Pike VM instead of being refused
            m.tdfas.push_back(placeholder_tdfa(m.programs[e], m.symbols));
            m.pike_only.push_back(1);
            m.pike[e] = build_pike_tables(m.programs[e], m.symbols);
        }
    }
    return m;
}

}  // namespace gorp
