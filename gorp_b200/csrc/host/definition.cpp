// See definition.hpp. Written from the reference's observable behaviour; structure is our own:
// one `Reader` object holds the three declaration tables and runs tokenise -> resolve -> flatten.
#include "definition.hpp"

#include <map>
#include <memory>
#include <set>
#include <unordered_map>

namespace gorp {
namespace {

using u16 = char16_t;

// ---------------------------------------------------------------- error context
struct LineCtx {
    std::string src;
    int row = 0;
    ustring text;
    [[noreturn]] void fail(size_t col, const std::string& msg) const {
        throw DefinitionParseError(strfmt("[%s (%d,%zu)]: %s", src.c_str(), row, col, msg.c_str()));
    }
};
using Ctx = std::shared_ptr<const LineCtx>;

std::string u8(const ustring& s) { return utf16_to_utf8(s); }

// ---------------------------------------------------------------- logical lines (InputLineReader.java:69-150)
std::vector<ustring> physical_lines(const ustring& t) {
    std::vector<ustring> out;
    size_t start = 0, i = 0, n = t.size();
    while (i < n) {
        u16 c = t[i];
        if (c == u'\n' || c == u'\r') {
            out.emplace_back(t.substr(start, i - start));
            if (c == u'\r' && i + 1 < n && t[i + 1] == u'\n') ++i;
            start = ++i;
        } else {
            ++i;
        }
    }
    if (start < n) out.emplace_back(t.substr(start));
    return out;
}

bool blank_or_comment(const ustring& l) {
    for (u16 c : l) {
        if (c <= 0x20) continue;
        return c == u'#';
    }
    return true;
}

struct LineSource {
    std::vector<ustring> lines;
    size_t pos = 0;
    int row = 0;
    std::string src;

    [[noreturn]] void fail(const std::string& msg) const {
        throw DefinitionParseError(strfmt("(%s, row %d): %s", src.c_str(), row, msg.c_str()));
    }
    // Returns false at end of input.
    bool next(Ctx& out) {
        const ustring* first = nullptr;
        while (pos < lines.size()) {
            const ustring& cand = lines[pos++];
            ++row;
            if (!blank_or_comment(cand)) { first = &cand; break; }
        }
        if (!first) return false;
        auto ctx = std::make_shared<LineCtx>();
        ctx->src = src;
        ctx->row = row;
        ustring acc = *first;
        while (!acc.empty() && acc.back() == u'\\') {
            acc.pop_back();
            // continuation lines are taken verbatim: no comment/blank filtering, no trimming
            if (pos >= lines.size()) fail("Unexpected end-of-input when expecting line continuation'");
            const ustring& seg = lines[pos++];
            ++row;
            if (!seg.empty() && seg.back() == u'\\') { acc += seg; continue; }  // loop pops the backslash
            acc += seg;
            break;
        }
        ctx->text = std::move(acc);
        out = ctx;
        return true;
    }
};

// ---------------------------------------------------------------- token helpers (TokenHelper.java)
inline bool ws(u16 c) { return c <= u' '; }
inline bool digit(u16 c) { return c >= u'0' && c <= u'9'; }
inline bool ident_start(u16 c) {
    if (c < 0x80) return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_' || c == '$';
    if (c >= 0xA0 && c <= 0xBF) return c == 0xAA || c == 0xB5 || c == 0xBA || (c >= 0xA2 && c <= 0xA5);
    if (c >= 0x2000 && c <= 0x206F) return false;
    if (c >= 0xD800 && c <= 0xDFFF) return false;
    return c >= 0xC0 && c != 0xD7 && c != 0xF7;
}
inline bool ident_part(u16 c) { return ident_start(c) || digit(c); }

struct Name {
    bool present = false;
    ustring text;
    size_t rest = 0;
};

Name parse_name(const char* kind, const Ctx& cx, size_t ix, bool allow_numbers) {
    const ustring& s = cx->text;
    const size_t end = s.size();
    if (ix >= end) cx->fail(end, strfmt("Missing %s name", kind));
    Name r;
    u16 c = s[ix];
    if (c == u'"' || c == u'\'') {
        size_t q = s.find(c, ix + 1);
        if (q == ustring::npos) cx->fail(end, strfmt("Missing closing quote ('%c') for %s name", static_cast<char>(c), kind));
        r.present = true;
        r.text = s.substr(ix + 1, q - ix - 1);
        r.rest = q + 1;
    } else if (!ident_start(c)) {
        if (digit(c)) {
            if (!allow_numbers)
                cx->fail(ix, strfmt("Invalid variable reference instead of %s name: can not use variable references here "
                                    "(missing parenthesis after template name?)", kind));
            size_t st = ix;
            while (ix < end && digit(s[ix])) ++ix;
            r.present = true;
            r.text = s.substr(st, ix - st);
        }
        r.rest = ix;
    } else {
        size_t st = ix++;
        while (ix < end && ident_part(s[ix])) ++ix;
        r.present = true;
        r.text = s.substr(st, ix - st);
        r.rest = ix;
    }
    return r;
}

Name parse_name_skip_space(const char* kind, const Ctx& cx, size_t ix) {
    Name r = parse_name(kind, cx, ix, false);
    const ustring& s = cx->text;
    if (r.rest >= s.size()) return r;
    if (!ws(s[r.rest])) cx->fail(r.rest, strfmt("Missing space character after %s name '%s'", kind, u8(r.text).c_str()));
    while (r.rest < s.size() && ws(s[r.rest])) ++r.rest;
    return r;
}

long nonneg_number(const ustring& s) {
    if (s.empty()) return -1;
    long v = 0;
    for (u16 c : s) {
        if (!digit(c)) return -1;
        v = v * 10 + (c - u'0');
        if (v > 100000000) return 100000000;
    }
    return v;
}

// index just past the single `ch` (whitespace allowed around), or npos when not found
size_t match_remaining(const ustring& s, size_t ix, u16 ch) {
    bool found = false;
    while (ix < s.size()) {
        u16 c = s[ix++];
        if (c == ch) {
            if (found) break;
            found = true;
        } else if (!ws(c)) {
            break;
        }
    }
    return found ? ix : ustring::npos;
}

// ---------------------------------------------------------------- pieces
enum class Kind { Text, Pattern, PatternRef, TemplateRef, TemplateParam, ExtractorParam, Extractor };

struct Piece;
using PiecePtr = std::shared_ptr<Piece>;
struct Piece {
    Kind kind;
    Ctx cx;
    size_t off = 0;
    ustring text;                  // literal text / pattern text / referenced name / extractor name
    int position = -1;             // parameter position (TemplateParam/ExtractorParam/positional Extractor)
    bool has_params = false;       // TemplateRef: parameter list present (even if empty list was never appended)
    std::vector<PiecePtr> kids;    // Extractor body or TemplateRef parameters
    [[noreturn]] void fail(const std::string& m) const { cx->fail(off, m); }
};

PiecePtr mk(Kind k, const Ctx& cx, size_t off, ustring text = {}) {
    auto p = std::make_shared<Piece>();
    p->kind = k;
    p->cx = cx;
    p->off = off;
    p->text = std::move(text);
    return p;
}

const char* kind_name(Kind k) {
    switch (k) {
        case Kind::Text: return "LiteralText";
        case Kind::Pattern: return "LiteralPattern";
        case Kind::PatternRef: return "PatternReference";
        case Kind::TemplateRef: return "TemplateReference";
        case Kind::TemplateParam: return "TemplateParameterReference";
        case Kind::ExtractorParam: return "ExtractorParameterReference";
        default: return "ExtractorExpression";
    }
}

struct ParamTypes {  // ParameterCollector
    std::string types;
    void add(const Ctx& cx, size_t off, long pos, char type) {
        size_t p = static_cast<size_t>(pos - 1);
        if (types.size() <= p) types.resize(p + 1, '\0');
        char old = types[p];
        if (old != type && old != '\0')
            cx->fail(off, strfmt("Inconsistent references to parameter %ld: %c vs %c", pos, old, type));
        types[p] = type;
    }
};

struct RawDef {  // a declared pattern / template / extraction template before resolution
    Ctx cx;
    ustring name;
    size_t body_start = 0;
    bool parametric = false;
    ParamTypes params;
    std::vector<PiecePtr> parts;
};

struct Cooked {
    ustring name;
    bool parametric = false;
    std::string types;
    std::vector<PiecePtr> parts;
};

template <class V>
struct Ordered {  // LinkedHashMap: insertion order, replace keeps position
    std::vector<std::pair<ustring, V>> items;
    std::unordered_map<std::u16string, size_t> idx;
    V* find(const ustring& k) {
        auto it = idx.find(k);
        return it == idx.end() ? nullptr : &items[it->second].second;
    }
    bool put(const ustring& k, V v) {  // returns true when a previous value was replaced
        auto it = idx.find(k);
        if (it != idx.end()) { items[it->second].second = std::move(v); return true; }
        idx.emplace(k, items.size());
        items.emplace_back(k, std::move(v));
        return false;
    }
};

struct RawExtraction {
    std::shared_ptr<RawDef> tmpl;
    std::vector<std::string> appends;
};

// ---------------------------------------------------------------- the reader
class Reader {
  public:
    Reader(const ustring& text, const std::string& src) {
        in_.lines = physical_lines(text);
        in_.src = src;
    }

    std::vector<ExtractionStrings> run() {
        read_declarations();
        if (extractions_.items.empty())
            throw DefinitionParseError("No extraction definitions found from definition");
        for (auto& kv : patterns_.items)
            if (!cooked_patterns_.count(kv.first)) cooked_patterns_[kv.first] = resolve_pattern(kv.first, *kv.second, nullptr);
        for (auto& kv : templates_.items) {
            if (cooked_templates_.count(kv.first)) continue;
            auto ct = std::make_shared<Cooked>();
            ct->name = kv.first;
            ct->parametric = kv.second->parametric;
            ct->types = kv.second->params.types;
            resolve_contents(true, kv.second->name, kv.second->parts, ct->parts, nullptr, kv.first);
            cooked_templates_[kv.first] = ct;
        }
        std::vector<ExtractionStrings> out;
        for (auto& kv : extractions_.items) {
            RawDef& raw = *kv.second.tmpl;
            std::vector<PiecePtr> cooked;
            resolve_contents(false, raw.name, raw.parts, cooked, nullptr, raw.name);
            ExtractionStrings xs;
            xs.name = kv.first;
            std::vector<PiecePtr> flat;
            flatten(cooked, flat, xs.extractor_names, nullptr, true);
            for (auto& p : flat) emit(*p, xs.automaton_regex, xs.jdk_regex);
            for (size_t i = 0; i < kv.second.appends.size(); ++i) {
                if (i) xs.append_json += "\n";
                xs.append_json += kv.second.appends[i];
            }
            out.push_back(std::move(xs));
        }
        return out;
    }

  private:
    LineSource in_;
    Ordered<std::shared_ptr<RawDef>> patterns_, templates_;
    Ordered<RawExtraction> extractions_;
    std::map<ustring, ustring> cooked_patterns_;
    std::map<ustring, std::shared_ptr<Cooked>> cooked_templates_;

    // ---- declarations (DefinitionReader.java:126-181, :189-207, :260-294, :528-594)
    void read_declarations() {
        Ctx cx;
        while (in_.next(cx)) {
            const ustring& s = cx->text;
            // keyword = \s*(\w*)\s*(.*)
            size_t i = 0;
            auto jws = [](u16 c) { return c == ' ' || (c >= 9 && c <= 13); };
            while (i < s.size() && jws(s[i])) ++i;
            size_t k0 = i;
            while (i < s.size() && s[i] < 0x80 && (ident_part(s[i]) && s[i] != '$')) ++i;
            ustring kw = s.substr(k0, i - k0);
            while (i < s.size() && jws(s[i])) ++i;
            if (kw == u"pattern") decl_pattern(cx, i);
            else if (kw == u"template") decl_template(cx, i);
            else if (kw == u"extract") decl_extraction(cx, i);
            else cx->fail(0, strfmt("Unrecognized keyword \"%s\" encountered; expected one of (pattern, template, extract)", u8(kw).c_str()));
        }
        for (auto& kv : patterns_.items) tokenize_pattern(*kv.second);
        for (auto& kv : templates_.items) {
            RawDef& t = *kv.second;
            tokenize_template(t.cx, t.body_start, t.parts, t.name, -1, "template '" + u8(t.name) + "' definition",
                              t.parametric ? &t.params : nullptr);
        }
        for (auto& kv : extractions_.items) {
            RawDef& t = *kv.second.tmpl;
            tokenize_template(t.cx, t.body_start, t.parts, t.name, 0, "extraction template for '" + u8(t.name) + "'", nullptr);
        }
    }

    static size_t type_marker(u16 m, const ustring& s, size_t ix) {
        for (; ix < s.size(); ++ix) {
            if (s[ix] == m) return ix;
            if (ws(s[ix])) break;
        }
        return ustring::npos;
    }

    void decl_pattern(const Ctx& cx, size_t off) {
        size_t ix = type_marker(u'%', cx->text, off);
        if (ix == ustring::npos) cx->fail(off, "Pattern name must be prefixed with '%'");
        Name n = parse_name_skip_space("pattern", cx, ix + 1);
        auto d = std::make_shared<RawDef>();
        d->cx = cx;
        d->name = n.text;
        d->body_start = n.rest;
        if (patterns_.find(n.text)) cx->fail(ix + 1, strfmt("Duplicate pattern definition for name '%s'", u8(n.text).c_str()));
        patterns_.put(n.text, d);
    }

    void decl_template(const Ctx& cx, size_t off) {
        const ustring& s = cx->text;
        size_t ix = type_marker(u'@', s, off);
        if (ix == ustring::npos) cx->fail(off, "Template name must be prefixed with '@'");
        ++ix;
        Name n = parse_name("template", cx, ix, false);
        size_t name_off = ix;
        ix = n.rest;
        auto d = std::make_shared<RawDef>();
        if (ix + 1 < s.size() && s[ix] == u'(' && s[ix + 1] == u')') {
            ix += 2;
            d->parametric = true;
        }
        size_t ix2 = ix;
        while (ix2 < s.size() && ws(s[ix2])) ++ix2;
        if (ix2 == ix) cx->fail(ix, strfmt("Missing space character after template name '%s'", u8(n.text).c_str()));
        d->cx = cx;
        d->name = n.text;
        d->body_start = ix2;
        if (templates_.find(n.text)) cx->fail(name_off, strfmt("Duplicate template definition for name '%s'", u8(n.text).c_str()));
        templates_.put(n.text, d);
    }

    void decl_extraction(Ctx cx, size_t off) {
        Name n = parse_name_skip_space("extraction", cx, off);
        std::string nm = u8(n.text);
        if (match_remaining(cx->text, n.rest, u'{') != cx->text.size())
            cx->fail(n.rest, strfmt("Unexpected content for extraction '%s': expected only opening '{'", nm.c_str()));
        RawExtraction x;
        size_t ix = 0;
        for (;;) {
            if (!in_.next(cx)) in_.fail(strfmt("Unexpected end-of-input in extraction '%s' definition", nm.c_str()));
            const ustring& s = cx->text;
            size_t close = match_remaining(s, 0, u'}');
            if (close != ustring::npos) {
                if (close >= s.size()) break;
                cx->fail(n.rest, strfmt("Unexpected content after closing '}' for extraction '%s'", nm.c_str()));
            }
            ix = 0;
            while (ix < s.size() && ws(s[ix])) ++ix;
            Name prop = parse_name_skip_space("extraction", cx, ix);
            ix = prop.rest;
            if (prop.text == u"template") {
                if (x.tmpl) cx->fail(ix, "More than one 'template' specified for '" + nm + "'");
                x.tmpl = std::make_shared<RawDef>();
                x.tmpl->cx = cx;
                x.tmpl->body_start = ix;
            } else if (prop.text == u"append") {
                std::string raw = u8(s.substr(ix));
                size_t a = 0, b = raw.size();
                while (a < b && static_cast<unsigned char>(raw[a]) <= ' ') ++a;
                while (b > a && static_cast<unsigned char>(raw[b - 1]) <= ' ') --b;
                raw = raw.substr(a, b - a);
                if (!raw.empty()) {
                    if (raw[0] != '{' && raw[0] == '"') raw = "{" + raw + "}";
                    if (raw.front() != '{' || raw.back() != '}')
                        cx->fail(ix, "Invalid 'append' value: must be JSON Object, or sequence of key/value pairs");
                    x.appends.push_back(raw);
                }
            } else {
                cx->fail(ix, strfmt("Unrecognized extraction property \"%s\" encountered; expected one of (template, append)",
                                    u8(prop.text).c_str()));
            }
        }
        if (!x.tmpl) cx->fail(ix, strfmt("Missing 'template' for extraction '%s'", nm.c_str()));
        extractions_.put(n.text, std::move(x));
    }

    // ---- pattern bodies (DefinitionReader.java:209-258)
    void tokenize_pattern(RawDef& d) {
        const Ctx& cx = d.cx;
        const ustring& s = cx->text;
        size_t ix = s.find(u'%', d.body_start);
        if (ix == ustring::npos) {
            d.parts.push_back(mk(Kind::Pattern, cx, d.body_start, s.substr(d.body_start)));
            return;
        }
        ustring lit = s.substr(d.body_start, ix - d.body_start);
        auto flush = [&] {
            if (!lit.empty()) d.parts.push_back(mk(Kind::Pattern, cx, d.body_start, lit));
            lit.clear();
        };
        while (ix < s.size()) {
            u16 c = s[ix++];
            if (c != u'%') { lit.push_back(c); continue; }
            if (ix == s.size()) cx->fail(ix, strfmt("Orphan '%%' at end of pattern '%s' definition", u8(d.name).c_str()));
            if (s[ix] == u'%') { lit.push_back(u'%'); ++ix; continue; }
            Name ref = parse_name("pattern", cx, ix, false);
            flush();
            d.parts.push_back(mk(Kind::PatternRef, cx, ix, ref.text));
            ix = ref.rest;
        }
        flush();
    }

    // ---- template / extractor bodies (DefinitionReader.java:303-392)
    size_t tokenize_template(const Ctx& cx, size_t ix, std::vector<PiecePtr>& out, const ustring& owner, int parens,
                             const std::string& desc, ParamTypes* vars) {
        const ustring& s = cx->text;
        const size_t end = s.size();
        ustring lit;
        size_t lit_start = ix;
        auto flush = [&] {
            if (!lit.empty()) out.push_back(mk(Kind::Text, cx, lit_start, lit));
            lit.clear();
        };
        while (ix < end) {
            u16 c = s[ix++];
            if (c == u'%' || c == u'@' || c == u'$') {
                if (ix == end) cx->fail(ix, strfmt("Orphan '%c' at end of %s", static_cast<char>(c), desc.c_str()));
                u16 d = s[ix];
                if (d == c) { lit.push_back(c); ++ix; continue; }
                flush();
                if (c == u'%') {
                    if (d == u'{') {
                        ++ix;
                        size_t st = ix, depth = 1, i = ix;
                        bool closed = false;
                        while (i < end) {  // TokenHelper.parseInlinePattern
                            u16 e = s[i++];
                            if (e == u'\\') { ++i; continue; }
                            if (e == u'{') ++depth;
                            else if (e == u'}' && --depth == 0) { closed = true; break; }
                        }
                        if (!closed) cx->fail(st, "Missing closing '{' for inline pattern");
                        out.push_back(mk(Kind::Pattern, cx, st, s.substr(st, i - 1 - st)));
                        ix = i;
                    } else {
                        Name n = parse_name("pattern", cx, ix, false);
                        out.push_back(mk(Kind::PatternRef, cx, ix, n.text));
                        ix = n.rest;
                    }
                } else if (c == u'@') {
                    ix = tokenize_template_ref(cx, ix, out, owner, desc, vars);
                } else {
                    Name n = parse_name("extractor", cx, ix, vars != nullptr);
                    ix = n.rest;
                    PiecePtr ex = mk(Kind::Extractor, cx, ix, n.text);
                    long pos = (vars && n.present) ? nonneg_number(n.text) : -1;
                    if (vars && pos >= 0) {
                        if (pos < 1 || pos > 999999) cx->fail(ix, strfmt("Invalid extractor name parameter %ld in %s", pos, desc.c_str()));
                        vars->add(cx, ix, pos, '$');
                        ex->position = static_cast<int>(pos);
                    }
                    out.push_back(ex);
                    if (ix >= end || s[ix] != u'(')
                        cx->fail(ix, strfmt("Invalid declaration for extractor '%s': missing opening parenthesis", u8(ex->text).c_str()));
                    ix = tokenize_template(cx, ix + 1, ex->kids, ex->text, 1, "extractor '" + u8(ex->text) + "' expression", vars);
                }
                lit_start = ix;
                continue;
            }
            if (parens > 0) {
                if (c == u'(') ++parens;
                else if (c == u')' && --parens == 0) break;
            }
            lit.push_back(c);
        }
        flush();
        if (parens > 0) cx->fail(ix, strfmt("Missing closing parenthesis at end of %s", desc.c_str()));
        return ix;
    }

    size_t tokenize_template_ref(const Ctx& cx, size_t ix, std::vector<PiecePtr>& out, const ustring& owner,
                                 const std::string& desc, ParamTypes* vars) {
        Name n = parse_name("template parameter", cx, ix, vars != nullptr);
        ix = n.rest;
        long pos = (vars && n.present) ? nonneg_number(n.text) : -1;
        if (vars && pos >= 0) {
            if (pos < 1 || pos > 999999) cx->fail(ix, strfmt("Invalid template parameter %ld in %s", pos, desc.c_str()));
            vars->add(cx, ix, pos, '@');
            PiecePtr p = mk(Kind::TemplateParam, cx, ix, owner);
            p->position = static_cast<int>(pos);
            out.push_back(p);
            return ix;
        }
        auto* target = n.present ? templates_.find(n.text) : nullptr;
        if (!target) cx->fail(ix, strfmt("Referencing non-existing template '@%s' from '%s'", u8(n.text).c_str(), desc.c_str()));
        PiecePtr ref = mk(Kind::TemplateRef, cx, ix, n.text);
        out.push_back(ref);
        if ((*target)->parametric) ix = tokenize_param_list(cx, ix, *ref, desc, vars);
        return ix;
    }

    size_t tokenize_param_list(const Ctx& cx, size_t ix, Piece& ref, const std::string& desc, ParamTypes* vars) {
        const ustring& s = cx->text;
        const size_t end = s.size();
        std::string rn = u8(ref.text);
        if (ix >= end || s[ix] != u'(') cx->fail(ix, strfmt("Missing parameter list for template reference '@%s'", rn.c_str()));
        ++ix;
        for (int idx = 1; ix < end; ++idx) {
            u16 c = s[ix++];
            if (c == u')') return ix;
            if (idx > 1) {
                if (c != u',')
                    cx->fail(ix, strfmt("Unexpected character '%c' in template parameter list for '@%s': expected either ',' or ')')'",
                                        static_cast<char>(c), rn.c_str()));
                if (ix >= end) break;
                c = s[ix++];
            }
            if (c == u'@') {
                size_t before = ref.kids.size();
                ix = tokenize_template_ref(cx, ix, ref.kids, ref.text, desc, vars);
                if (ref.kids.size() > before) ref.has_params = true;
            } else if (c == u'$') {
                Name n = parse_name("extractor parameter", cx, ix, vars != nullptr);
                ix = n.rest;
                long pos = (vars && n.present) ? nonneg_number(n.text) : -1;
                PiecePtr p;
                if (vars && pos >= 0) {
                    if (pos < 1 || pos > 999999) cx->fail(ix, strfmt("Invalid extractor parameter %ld in %s", pos, desc.c_str()));
                    vars->add(cx, ix, pos, '$');
                    p = mk(Kind::ExtractorParam, cx, ix, ref.text);
                    p->position = static_cast<int>(pos);
                } else {
                    p = mk(Kind::Extractor, cx, ix, n.text);
                }
                ref.kids.push_back(p);
                ref.has_params = true;
            } else {
                cx->fail(ix, strfmt("Unexpected character '%c' in template parameter list for '@%s': expected either type marker "
                                    "'@' or closing ')'", static_cast<char>(c), rn.c_str()));
            }
        }
        cx->fail(ix, strfmt("Unexpected end of line within parameter list for template '@%s'", rn.c_str()));
    }

    // ---- pattern resolution: pure string splicing with cycle check (CookedDefinitions.java:57-132)
    static std::string chain(const char* marker, const std::vector<ustring>* stack, const ustring& last) {
        if (!stack) return "";
        std::string r = "(";
        for (auto& s : *stack) r += marker + u8(s) + "->";
        return r + marker + u8(last) + ")";
    }

    ustring resolve_pattern(const ustring& name, RawDef& d, std::vector<ustring>* stack) {
        std::vector<ustring> local;
        ustring out;
        for (auto& p : d.parts) {
            if (p->kind == Kind::Pattern) { out += p->text; continue; }
            if (!stack) stack = &local;
            const ustring& to = p->text;
            auto hit = cooked_patterns_.find(to);
            if (hit != cooked_patterns_.end()) { out += hit->second; continue; }
            stack->push_back(name);
            for (auto& s : *stack)
                if (s == to) p->fail(strfmt("Cyclic pattern reference to '%%%s' %s", u8(to).c_str(), chain("%", stack, to).c_str()));
            auto* raw = patterns_.find(to);
            if (!raw) p->fail(strfmt("Referencing non-existing pattern '%%%s' %s", u8(to).c_str(), chain("%", stack, to).c_str()));
            ustring sub = resolve_pattern(to, **raw, stack);
            cooked_patterns_[to] = sub;
            stack->pop_back();
            out += sub;
        }
        return out;
    }

    // ---- template resolution (CookedDefinitions.java:144-242)
    // `declared`: true while resolving declared templates (forward references can still be unresolved).
    void resolve_contents(bool declared, const ustring& name, const std::vector<PiecePtr>& in, std::vector<PiecePtr>& out,
                          std::vector<ustring>* stack, const ustring& top) {
        std::vector<ustring> local;
        for (auto& d : in) {
            switch (d->kind) {
                case Kind::Text:
                case Kind::Pattern:
                    out.push_back(d);
                    break;
                case Kind::PatternRef: {
                    auto hit = cooked_patterns_.find(d->text);
                    if (hit == cooked_patterns_.end())
                        d->fail(strfmt("Referencing non-existing pattern '%%%s' from template '%s' %s", u8(d->text).c_str(),
                                       u8(top).c_str(), chain("@", stack, name).c_str()));
                    out.push_back(mk(Kind::Pattern, d->cx, d->off, hit->second));
                    break;
                }
                case Kind::TemplateRef: {
                    if (d->has_params) { out.push_back(d); break; }
                    if (!stack) stack = &local;
                    auto t = resolve_template_ref(declared, name, *d, stack, top);
                    for (auto& p : t->parts) out.push_back(p);
                    break;
                }
                case Kind::Extractor: {
                    PiecePtr r = std::make_shared<Piece>(*d);
                    r->kids.clear();
                    if (!stack) stack = &local;
                    resolve_contents(declared, name, d->kids, r->kids, stack, top);
                    out.push_back(r);
                    break;
                }
                case Kind::TemplateParam:
                    out.push_back(d);
                    break;
                default:
                    d->cx->fail(0, strfmt("Internal error: unexpected definition type %s when resolving template definition '%s'",
                                          kind_name(d->kind), u8(top).c_str()));
            }
        }
    }

    std::shared_ptr<Cooked> resolve_template_ref(bool declared, const ustring& from, Piece& ref, std::vector<ustring>* stack,
                                                 const ustring& top) {
        (void)top;
        const ustring& to = ref.text;
        auto hit = cooked_templates_.find(to);
        if (hit != cooked_templates_.end()) return hit->second;
        stack->push_back(from);
        for (auto& s : *stack)
            if (s == to) ref.fail(strfmt("Cyclic template reference to '%%%s' %s", u8(to).c_str(), chain("@", stack, to).c_str()));
        auto* raw = declared ? templates_.find(to) : nullptr;
        if (!raw) ref.fail(strfmt("Referencing non-existing template '%%%s' %s", u8(to).c_str(), chain("@", stack, to).c_str()));
        // Reference behaviour kept on purpose: a template that is referenced before its own declaration was
        // resolved is registered with NO contents (the reference resolves the new template's own empty part
        // list, CookedDefinitions.java:235-237) and stays empty for every later use.
        auto ct = std::make_shared<Cooked>();
        ct->name = to;
        ct->parametric = (*raw)->parametric;
        ct->types = (*raw)->params.types;
        cooked_templates_[to] = ct;
        stack->pop_back();
        return ct;
    }

    // ---- extraction flattening (CookedDefinitions.java:255-453)
    using Bindings = std::vector<PiecePtr>;

    void flatten(const std::vector<PiecePtr>& in, std::vector<PiecePtr>& out, std::vector<ustring>& names, const Bindings* bind,
                 bool top = false) {
        for (PiecePtr part : in) {
            if (part->kind == Kind::TemplateParam) {
                if (top) part->fail(strfmt("Internal error: should not encounter template parameter %s#%d", u8(part->text).c_str(), part->position));
                if (!bind) part->fail(strfmt("Invalid parameter variable reference @%d; template takes no parameters", part->position));
                if (part->position < 1 || static_cast<size_t>(part->position) > bind->size())
                    part->fail(strfmt("Invalid parameter variable reference @%d; template takes %zu parameters", part->position, bind->size()));
                part = (*bind)[part->position - 1];
            }
            switch (part->kind) {
                case Kind::Text:
                case Kind::Pattern:
                    out.push_back(part);
                    break;
                case Kind::PatternRef: {
                    auto hit = cooked_patterns_.find(part->text);
                    if (hit == cooked_patterns_.end())
                        throw DefinitionParseError(strfmt("Internal error: non-existing pattern '%%%s', should have been caught earlier",
                                                          u8(part->text).c_str()));
                    out.push_back(mk(Kind::Pattern, part->cx, part->off, hit->second));
                    break;
                }
                case Kind::Extractor: {
                    ustring nm = part->text;
                    if (part->position >= 0) {
                        PiecePtr b = (bind && static_cast<size_t>(part->position) <= bind->size() && part->position >= 1)
                                         ? (*bind)[part->position - 1] : nullptr;
                        if (!b || b->kind != Kind::Extractor)
                            part->fail(strfmt("Internal error: unexpected extractor parameter of type %s (expecting ExtractorExpression)",
                                              b ? kind_name(b->kind) : "null"));
                        if (b->position >= 0)
                            part->fail(strfmt("Internal error: positional extractor parameter (%d) resolves to another positional (%d)",
                                              part->position, b->position));
                        nm = b->text;
                    }
                    for (auto& n : names)
                        if (n == nm) part->fail(strfmt("Duplicate extractor name ($%s)", u8(nm).c_str()));
                    names.push_back(nm);
                    PiecePtr r = mk(Kind::Extractor, part->cx, part->off, nm);
                    flatten(part->kids, r->kids, names, bind);
                    out.push_back(r);
                    break;
                }
                case Kind::TemplateRef:
                    expand_template(*part, out, names, bind);
                    break;
                default:
                    part->fail(strfmt("Internal error: unrecognized DefPiece %s", kind_name(part->kind)));
            }
        }
    }

    void expand_template(Piece& ref, std::vector<PiecePtr>& out, std::vector<ustring>& names, const Bindings* incoming) {
        auto hit = cooked_templates_.find(ref.text);
        if (hit == cooked_templates_.end()) ref.fail(strfmt("Internal error: reference to unknown template '@%s'", u8(ref.text).c_str()));
        Cooked& t = *hit->second;
        Bindings bound;
        const Bindings* use = nullptr;
        if (t.parametric) {
            size_t want = t.types.size();
            if (ref.kids.size() != want)
                ref.fail(strfmt("Parameter mismatch: template '@%s' expects %zu parameters; %zu passed", u8(ref.text).c_str(), want,
                                ref.kids.size()));
            for (size_t i = 0; i < want; ++i) {
                const PiecePtr& a = ref.kids[i];
                char exp = t.types[i];
                bool ok = exp == '@' ? a->kind == Kind::TemplateRef : exp == '$' ? a->kind == Kind::Extractor : false;
                if (exp != '@' && exp != '$')
                    throw DefinitionParseError(strfmt("Internal error: unrecognized template parameter type '%c'", exp));
                if (!ok)
                    ref.fail(strfmt("Parameter mismatch: template '@%s' expects type '%c' parameter, got %s", u8(ref.text).c_str(), exp,
                                    kind_name(a->kind)));
                bound.push_back(bind_argument(a, incoming));
            }
            use = &bound;
        }
        flatten(t.parts, out, names, use);
    }

    PiecePtr bind_argument(const PiecePtr& a, const Bindings* incoming) {  // _resolveParameters
        switch (a->kind) {
            case Kind::TemplateParam: {
                if (!incoming || a->position < 1 || static_cast<size_t>(a->position) > incoming->size())
                    a->fail(strfmt("Invalid parameter variable reference @%d; template has %zu parameters", a->position,
                                   incoming ? incoming->size() : size_t(0)));
                return (*incoming)[a->position - 1];
            }
            case Kind::TemplateRef: {
                if (!a->has_params) return a;
                PiecePtr r = std::make_shared<Piece>(*a);
                for (auto& k : r->kids) k = bind_argument(k, incoming);
                return r;
            }
            case Kind::Extractor: {
                PiecePtr r = std::make_shared<Piece>(*a);
                for (auto& k : r->kids) k = bind_argument(k, incoming);
                return r;
            }
            default:
                a->fail(strfmt("Internal error: unexpected template parameter type %s", kind_name(a->kind)));
        }
    }

    // ---- pieces -> the two strings (Gorp.java:94-129)
    void emit(const Piece& p, ustring& autom, ustring& jdk) {
        switch (p.kind) {
            case Kind::Pattern:
                try {
                    autom += massage_regexp_for_automaton(p.text);
                    jdk += massage_regexp_for_jdk(p.text);
                } catch (const std::invalid_argument& e) {
                    p.fail(strfmt("Invalid pattern definition, problem (java.lang.IllegalArgumentException): %s", e.what()));
                }
                break;
            case Kind::Text: {
                ustring q = quote_literal_as_regexp(p.text);
                autom += q;
                jdk += q;
                break;
            }
            case Kind::Extractor:
                autom.push_back(u'(');
                jdk.push_back(u'(');
                for (auto& k : p.kids) emit(*k, autom, jdk);
                autom.push_back(u')');
                jdk.push_back(u')');
                break;
            default:
                p.fail(strfmt("Unrecognized DefPiece in FlattenedExtraction: %s", kind_name(p.kind)));
        }
    }
};

}  // namespace

// ---------------------------------------------------------------- RegexHelper
ustring quote_literal_as_regexp(const ustring& t) {
    static const ustring special = u"()[]\\{}|*?+$^<>\"&";
    ustring out;
    for (size_t i = 0; i < t.size();) {
        char16_t c = t[i++];
        if (c == u' ' || c == u'\t') {
            while (i < t.size() && t[i] <= u' ') ++i;
            out += u"[ \t]+";  // literal TAB inside the class
        } else if (c == u'.') {
            out += u"\\.";
        } else if (special.find(c) != ustring::npos) {
            out.push_back(u'\\');
            out.push_back(c);
        } else {
            out.push_back(c);
        }
    }
    return out;
}

static bool alpha_or_digit(char16_t d) {
    if (d < 0x80) return (d >= '0' && d <= '9') || (d >= 'a' && d <= 'z') || (d >= 'A' && d <= 'Z');
    return ident_start(d);  // Character.isAlphabetic for the non-ASCII letters we can classify
}

ustring massage_regexp_for_automaton(const ustring& p) {
    if (p.find(u'\\') == ustring::npos) return p;
    static const ustring D = u"0-9", S = u" \b\f\n\r\t", W = u"a-zA-Z_0-9";
    ustring out;
    int depth = 0;
    for (size_t i = 0; i < p.size();) {
        char16_t c = p[i++];
        if (c == u'[') { out.push_back(c); ++depth; continue; }
        if (c == u']') { out.push_back(c); --depth; continue; }
        if (c != u'\\' || i >= p.size()) { out.push_back(c); continue; }
        bool after_open = depth > 0 && p[i - 2] == u'[';
        char16_t d = p[i++];
        const ustring* cls = nullptr;
        bool neg = false;
        switch (d) {
            case u'\\': break;
            case u'b': d = u'\b'; break;
            case u'f': d = u'\f'; break;
            case u'n': d = u'\n'; break;
            case u'r': d = u'\r'; break;
            case u't': d = u'\t'; break;
            case u'd': cls = &D; break;
            case u'D': cls = &D; neg = true; break;
            case u's': cls = &S; break;
            case u'S': cls = &S; neg = true; break;
            case u'w': cls = &W; break;
            case u'W': cls = &W; neg = true; break;
            default:
                if (alpha_or_digit(d))
                    throw std::invalid_argument(strfmt(
                        "Unrecognized backslash escape '\\%s; can only escape backslash (\\\\), use known control-codes (\\n, \\r, \\t), "
                        "escape non-alphanumeric (\\$, \\(, ...) or refer to a 'well-known' character class (\\s, \\S, \\d, \\D, \\w, \\W)",
                        utf16_to_utf8(ustring(1, d)).c_str()));
        }
        if (cls) {
            if (depth == 0) {
                out.push_back(u'[');
                if (neg) out.push_back(u'^');
                out += *cls;
                out.push_back(u']');
            } else {
                if (neg && !after_open)
                    throw std::invalid_argument(strfmt("Can not use negated character class \\%c within character class in position "
                                                       "other than first (Automaton limitation)", static_cast<char>(d)));
                if (neg) out.push_back(u'^');
                out += *cls;
            }
            continue;
        }
        out.push_back(c);
        out.push_back(d);
    }
    return out;
}

ustring massage_regexp_for_jdk(const ustring& p) {
    ustring out;
    for (size_t i = 0; i < p.size();) {
        char16_t c = p[i++];
        if (c == u'\\') {
            out.push_back(c);
            if (i < p.size()) out.push_back(p[i++]);
        } else if (c == u'(') {
            out += u"(?:";
        } else {
            out.push_back(c);
        }
    }
    return out;
}

std::vector<ExtractionStrings> read_definition(const ustring& text, const std::string& source_ref) {
    return Reader(text, source_ref).run();
}

}  // namespace gorp
