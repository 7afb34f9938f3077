#include "capture.hpp"

#include <algorithm>
#include <map>
#include <memory>
#include <set>
#include <unordered_map>

namespace gorp {
namespace {

constexpr uint32_t MAXCP = 0x10FFFF;
constexpr size_t MAX_INSTS = 30000;

Ranges normalize(Ranges r) {
    Ranges in;
    for (auto& p : r)
        if (p.first <= p.second) in.push_back(p);
    std::sort(in.begin(), in.end());
    Ranges out;
    for (auto& p : in) {
        if (!out.empty() && p.first <= out.back().second + 1) out.back().second = std::max(out.back().second, p.second);
        else out.push_back(p);
    }
    return out;
}

Ranges complement(const Ranges& r) {
    Ranges out;
    uint32_t prev = 0;
    for (auto& p : normalize(r)) {
        if (p.first > prev) out.push_back({prev, p.first - 1});
        prev = p.second + 1;
    }
    if (prev <= MAXCP) out.push_back({prev, MAXCP});
    return out;
}

const Ranges kDigit{{'0', '9'}};
const Ranges kSpace{{0x09, 0x0D}, {0x20, 0x20}};  // [ \t\n\x0B\f\r]
const Ranges kWord{{'0', '9'}, {'A', 'Z'}, {'_', '_'}, {'a', 'z'}};
const Ranges kLineTerm{{0x0A, 0x0A}, {0x0D, 0x0D}, {0x85, 0x85}, {0x2028, 0x2029}};

// ------------------------------------------------------------------ java.util.regex subset -> AST
struct RNode;
using RP = std::unique_ptr<RNode>;
struct RNode {
    enum K { Char, Set, Cat, Alt, Group, Rep } k = Cat;
    uint32_t ch = 0;
    Ranges cps;           // Set: code points
    bool by_codepoint = false;  // Set: CharProperty (reads a code point) vs BmpCharProperty (reads a unit)
    int group = 0;        // Group: 0 = non-capturing
    int mn = 0, mx = -1;
    bool greedy = true;
    std::vector<RP> kids;
};

bool nullable(const RNode& n) {
    switch (n.k) {
        case RNode::Char:
        case RNode::Set: return false;
        case RNode::Cat:
            for (auto& k : n.kids)
                if (!nullable(*k)) return false;
            return true;
        case RNode::Alt:
            for (auto& k : n.kids)
                if (nullable(*k)) return true;
            return false;
        case RNode::Group: return nullable(*n.kids[0]);
        case RNode::Rep: return n.mn == 0 || nullable(*n.kids[0]);
    }
    return false;
}

class JdkParser {
  public:
    explicit JdkParser(const ustring& s) : s_(s) {
        for (char16_t c : s)
            if (c >= 0xD800 && c <= 0xDFFF) throw UnsupportedError("surrogate / supplementary character inside a pattern");
    }
    RP parse() {
        RP e = alt();
        if (i_ < s_.size()) throw DefinitionParseError(strfmt("Unmatched closing ')' near index %zu", i_));
        return e;
    }
    int groups() const { return groups_; }

  private:
    const ustring& s_;
    size_t i_ = 0;
    int groups_ = 0;

    int peek(size_t k = 0) const { return i_ + k < s_.size() ? s_[i_ + k] : -1; }
    static RP node(RNode::K k) {
        auto n = std::make_unique<RNode>();
        n->k = k;
        return n;
    }
    static RP chr(uint32_t c) {
        RP n = node(RNode::Char);
        n->ch = c;
        return n;
    }
    static RP set(Ranges r, bool by_cp) {
        RP n = node(RNode::Set);
        n->cps = normalize(std::move(r));
        n->by_codepoint = by_cp;
        return n;
    }

    RP alt() {
        std::vector<RP> items;
        items.push_back(seq());
        while (peek() == '|') {
            ++i_;
            items.push_back(seq());
        }
        if (items.size() == 1) return std::move(items[0]);
        RP n = node(RNode::Alt);
        n->kids = std::move(items);
        return n;
    }
    RP seq() {
        RP n = node(RNode::Cat);
        for (;;) {
            int c = peek();
            if (c < 0 || c == '|' || c == ')') break;
            n->kids.push_back(closure(atom()));
        }
        return n;
    }
    RP atom() {
        int c = peek();
        switch (c) {
            case '(': {
                ++i_;
                RP g = node(RNode::Group);
                if (peek() == '?') {
                    if (peek(1) != ':') throw UnsupportedError("inline construct '(?' other than '(?:'");
                    i_ += 2;
                } else {
                    g->group = ++groups_;
                }
                g->kids.push_back(alt());
                if (peek() != ')') throw DefinitionParseError(strfmt("Unclosed group near index %zu", i_));
                ++i_;
                return g;
            }
            case '[': ++i_; return clazz();
            case '.': ++i_; return set(complement(kLineTerm), true);
            case '\\': ++i_; return escape();
            case '*': case '+': case '?':
                throw DefinitionParseError(strfmt("Dangling meta character '%c' near index %zu", c, i_));
            case '{': throw UnsupportedError("unescaped '{' at the start of an expression");
            case '^': case '$':
                throw UnsupportedError("'^'/'$' (literal for the DFA, anchor for java.util.regex)");
            default: ++i_; return chr(static_cast<uint32_t>(c));
        }
    }
    // returns true and fills `r`/`neg` for a predefined class, false + `c` for a single char
    bool escaped(bool in_class, Ranges& r, bool& neg, uint32_t& c) {
        int d = peek();
        if (d < 0) throw DefinitionParseError(in_class ? "Unclosed character class" : "Unexpected internal error (trailing backslash)");
        ++i_;
        neg = false;
        switch (d) {
            case 'd': r = kDigit; return true;
            case 'D': r = kDigit; neg = true; return true;
            case 's': r = kSpace; return true;
            case 'S': r = kSpace; neg = true; return true;
            case 'w': r = kWord; return true;
            case 'W': r = kWord; neg = true; return true;
            case 't': c = 0x09; return false;
            case 'n': c = 0x0A; return false;
            case 'r': c = 0x0D; return false;
            case 'f': c = 0x0C; return false;
            case 'b':
                if (in_class) throw DefinitionParseError("Illegal/unsupported escape sequence \\b inside a character class");
                throw UnsupportedError("\\b (backspace for the DFA, word boundary for java.util.regex)");
            default:
                if ((d >= '0' && d <= '9') || (d >= 'a' && d <= 'z') || (d >= 'A' && d <= 'Z') || d >= 0x80)
                    throw UnsupportedError(strfmt("escape sequence \\%s", utf16_to_utf8(ustring(1, static_cast<char16_t>(d))).c_str()));
                c = static_cast<uint32_t>(d);
                return false;
        }
    }
    RP escape() {
        Ranges r;
        bool neg;
        uint32_t c;
        if (escaped(false, r, neg, c)) return neg ? set(complement(r), true) : set(r, false);
        return chr(c);
    }
    RP clazz() {
        bool negate = false;
        if (peek() == '^') { negate = true; ++i_; }
        Ranges acc;
        bool have = false, bits_only = true;
        for (;;) {
            int c = peek();
            if (c < 0) throw DefinitionParseError(strfmt("Unclosed character class near index %zu", i_));
            if (c == ']' && have) { ++i_; break; }
            if (c == '[') throw UnsupportedError("nested character class");
            if (c == '&' && peek(1) == '&') throw UnsupportedError("character class intersection '&&'");
            uint32_t lo;
            if (c == '\\') {
                ++i_;
                Ranges r;
                bool neg;
                if (escaped(true, r, neg, lo)) {
                    if (neg) { r = complement(r); bits_only = false; }
                    acc.insert(acc.end(), r.begin(), r.end());
                    have = true;
                    continue;
                }
            } else {
                ++i_;
                lo = static_cast<uint32_t>(c);
            }
            have = true;
            if (peek() == '-') {
                int e = peek(1);
                if (e == '[') throw UnsupportedError("nested character class");
                if (e >= 0 && e != ']') {
                    ++i_;
                    uint32_t hi;
                    if (peek() == '\\') {
                        ++i_;
                        Ranges r;
                        bool neg;
                        if (escaped(true, r, neg, hi)) throw DefinitionParseError(strfmt("Illegal character range near index %zu", i_));
                    } else {
                        hi = static_cast<uint32_t>(peek());
                        ++i_;
                    }
                    if (hi < lo) throw DefinitionParseError(strfmt("Illegal character range near index %zu", i_));
                    acc.push_back({lo, hi});
                    bits_only = false;
                    continue;
                }
            }
            acc.push_back({lo, lo});
            if (lo >= 256) bits_only = false;
        }
        if (negate) return set(complement(acc), true);
        return set(acc, !bits_only);
    }
    RP closure(RP a) {
        int c = peek();
        int mn, mx;
        if (c == '?') { ++i_; mn = 0; mx = 1; }
        else if (c == '*') { ++i_; mn = 0; mx = -1; }
        else if (c == '+') { ++i_; mn = 1; mx = -1; }
        else if (c == '{') {
            size_t j = i_ + 1, st = j;
            auto num = [&](size_t a0, size_t b0) {
                if (b0 - a0 > 6) throw UnsupportedError("repetition count too large");
                int v = 0;
                for (size_t k = a0; k < b0; ++k) v = v * 10 + (s_[k] - u'0');
                return v;
            };
            while (j < s_.size() && s_[j] >= u'0' && s_[j] <= u'9') ++j;
            if (j == st) throw DefinitionParseError(strfmt("Illegal repetition near index %zu", i_));
            mn = num(st, j);
            mx = mn;
            if (j < s_.size() && s_[j] == u',') {
                st = ++j;
                while (j < s_.size() && s_[j] >= u'0' && s_[j] <= u'9') ++j;
                mx = j > st ? num(st, j) : -1;
                if (mx != -1 && mx < mn) throw DefinitionParseError(strfmt("Illegal repetition range near index %zu", i_));
            }
            if (j >= s_.size() || s_[j] != u'}') throw DefinitionParseError(strfmt("Unclosed counted closure near index %zu", i_));
            i_ = j + 1;
        } else {
            return a;
        }
        bool greedy = true;
        if (peek() == '?') { ++i_; greedy = false; }
        else if (peek() == '+') throw UnsupportedError("possessive quantifier");
        int n = peek();
        if (n == '*' || n == '+' || n == '?' || n == '{') throw UnsupportedError("stacked quantifier");
        if ((mx == -1 || mx > 1) && nullable(*a)) throw UnsupportedError("quantified sub-expression can match the empty string");
        RP r = node(RNode::Rep);
        r->mn = mn;
        r->mx = mx;
        r->greedy = greedy;
        r->kids.push_back(std::move(a));
        return r;
    }
};

// ------------------------------------------------------------------ AST -> Pike program
class Emitter {
  public:
    CaptureProgram prog;
    void emit(const RNode& n) {
        auto& I = prog.insts;
        if (I.size() > MAX_INSTS) throw UnsupportedError("capture program too large (counted repetition unrolls too far)");
        switch (n.k) {
            case RNode::Char: I.push_back({OP_CHAR, static_cast<int32_t>(n.ch), 0}); break;
            case RNode::Set: emit_set(n); break;
            case RNode::Cat:
                for (auto& k : n.kids) emit(*k);
                break;
            case RNode::Alt: {
                std::vector<size_t> jumps;
                for (size_t k = 0; k < n.kids.size(); ++k) {
                    if (k + 1 < n.kids.size()) {
                        size_t sp = I.size();
                        I.push_back({OP_SPLIT, 0, 0});
                        emit(*n.kids[k]);
                        jumps.push_back(I.size());
                        I.push_back({OP_JMP, 0, 0});
                        I[sp].a = static_cast<int32_t>(sp + 1);
                        I[sp].b = static_cast<int32_t>(I.size());
                    } else {
                        emit(*n.kids[k]);
                    }
                }
                for (size_t j : jumps) I[j].a = static_cast<int32_t>(I.size());
                break;
            }
            case RNode::Group:
                if (n.group == 0) { emit(*n.kids[0]); break; }
                I.push_back({OP_SAVE, 2 * (n.group - 1), 0});
                emit(*n.kids[0]);
                I.push_back({OP_SAVE, 2 * (n.group - 1) + 1, 0});
                break;
            case RNode::Rep: {
                const RNode& body = *n.kids[0];
                for (int k = 0; k < n.mn; ++k) emit(body);
                auto split = [&](size_t sp, size_t end) {
                    if (n.greedy) { I[sp].a = static_cast<int32_t>(sp + 1); I[sp].b = static_cast<int32_t>(end); }
                    else { I[sp].a = static_cast<int32_t>(end); I[sp].b = static_cast<int32_t>(sp + 1); }
                };
                if (n.mx == -1) {
                    size_t sp = I.size();
                    I.push_back({OP_SPLIT, 0, 0});
                    emit(body);
                    I.push_back({OP_JMP, static_cast<int32_t>(sp), 0});
                    split(sp, I.size());
                } else {
                    std::vector<size_t> sps;
                    for (int k = n.mn; k < n.mx; ++k) {
                        sps.push_back(I.size());
                        I.push_back({OP_SPLIT, 0, 0});
                        emit(body);
                    }
                    for (size_t sp : sps) split(sp, I.size());
                }
                break;
            }
        }
    }

  private:
    std::map<Ranges, int> set_ids_;
    int set_id(const Ranges& r) {
        auto it = set_ids_.find(r);
        if (it != set_ids_.end()) return it->second;
        int id = static_cast<int>(prog.sets.size());
        prog.sets.push_back(r);
        set_ids_.emplace(r, id);
        return id;
    }
    void emit_set(const RNode& n) {
        auto& I = prog.insts;
        Ranges bmp, supp;
        for (auto& p : n.cps) {
            if (p.first <= 0xFFFF) bmp.push_back({p.first, std::min<uint32_t>(p.second, 0xFFFF)});
            if (p.second >= 0x10000) supp.push_back({std::max<uint32_t>(p.first, 0x10000), p.second});
        }
        bool pairs = false;
        if (n.by_codepoint && !supp.empty()) {
            if (!(supp.size() == 1 && supp[0].first == 0x10000 && supp[0].second == MAXCP))
                throw UnsupportedError("character class that contains only part of the supplementary planes");
            pairs = true;
        }
        if (!pairs) {
            I.push_back({OP_SET, set_id(bmp), 0});
        } else if (bmp.empty()) {
            I.push_back({OP_PAIRHI, 0, 0});
            I.push_back({OP_ANY, 0, 0});
        } else {
            // the two alternatives are disjoint on the input symbol, so their order carries no preference
            size_t sp = I.size();
            I.push_back({OP_SPLIT, static_cast<int32_t>(sp + 1), static_cast<int32_t>(sp + 3)});
            I.push_back({OP_SET, set_id(bmp), 0});
            I.push_back({OP_JMP, static_cast<int32_t>(sp + 5), 0});
            I.push_back({OP_PAIRHI, 0, 0});
            I.push_back({OP_ANY, 0, 0});
        }
    }
};

bool consuming(uint8_t op) { return op <= OP_PAIRHI; }

}  // namespace

CaptureProgram compile_jdk_regex(const ustring& regex) {
    JdkParser parser(regex);
    RP ast = parser.parse();
    if (parser.groups() > 31) throw UnsupportedError("more than 31 capturing groups in one extraction");
    Emitter em;
    em.emit(*ast);
    em.prog.insts.push_back({OP_MATCH, 0, 0});
    em.prog.n_groups = parser.groups();
    return std::move(em.prog);
}

SymbolClasses build_symbol_classes(const std::vector<CaptureProgram>& progs) {
    std::set<uint32_t> cuts{0};
    auto add = [&](uint32_t lo, uint32_t hi) {
        cuts.insert(lo);
        if (hi < 0xFFFF) cuts.insert(hi + 1);
    };
    for (auto& p : progs) {
        for (auto& in : p.insts)
            if (in.op == OP_CHAR) add(static_cast<uint32_t>(in.a), static_cast<uint32_t>(in.a));
        for (auto& s : p.sets)
            for (auto& r : s) add(r.first, r.second);
    }
    // merge intervals that no instruction distinguishes: signature = membership in every char/set
    std::vector<uint32_t> pts(cuts.begin(), cuts.end());
    std::map<std::vector<uint32_t>, uint32_t> sigs;
    std::vector<uint32_t> cls_of_pt(pts.size());
    // collect distinct predicates
    std::set<uint32_t> chars;
    std::set<Ranges> sets;
    for (auto& p : progs) {
        for (auto& in : p.insts)
            if (in.op == OP_CHAR) chars.insert(static_cast<uint32_t>(in.a));
        for (auto& s : p.sets) sets.insert(s);
    }
    for (size_t k = 0; k < pts.size(); ++k) {
        uint32_t c = pts[k];
        std::vector<uint32_t> sig;
        uint32_t idx = 0;
        for (uint32_t ch : chars) {
            if (ch == c) sig.push_back(idx);
            ++idx;
        }
        for (auto& s : sets) {
            auto it = std::upper_bound(s.begin(), s.end(), std::make_pair(c, 0xFFFFFFFFu));
            if (it != s.begin() && (it - 1)->second >= c) sig.push_back(idx);
            ++idx;
        }
        cls_of_pt[k] = sigs.emplace(std::move(sig), static_cast<uint32_t>(sigs.size())).first->second;
    }
    SymbolClasses sc;
    sc.classmap.resize(65536);
    size_t k = 0;
    for (uint32_t ch = 0; ch < 65536; ++ch) {
        if (k + 1 < pts.size() && ch == pts[k + 1]) ++k;
        sc.classmap[ch] = static_cast<uint16_t>(cls_of_pt[k]);
    }
    sc.pair_hi_class = static_cast<uint32_t>(sigs.size());
    sc.n_classes = sc.pair_hi_class + 1;
    return sc;
}

// ------------------------------------------------------------------ determinisation with tag registers
namespace {

struct Item {
    int pc;
    std::vector<int16_t> reg;  // per slot: register id, -1 = unset
};

struct ClosureEntry {
    int target;     // consuming instruction index, or -1 for MATCH
    uint64_t mask;  // SAVE slots passed on the way (set to the current position)
};

class TdfaBuilder {
  public:
    TdfaBuilder(const CaptureProgram& p, const SymbolClasses& sc, size_t max_states, size_t max_regs)
        : p_(p), sc_(sc), max_states_(max_states), max_regs_(max_regs), n_slots_(2 * p.n_groups) {
        const size_t C = sc.n_classes;
        // representative unit per class
        std::vector<int64_t> rep(C, -1);
        for (uint32_t ch = 0; ch < 65536; ++ch)
            if (rep[sc.classmap[ch]] < 0) rep[sc.classmap[ch]] = ch;
        accepts_.assign(p.insts.size(), std::vector<uint8_t>(C, 0));
        for (size_t i = 0; i < p.insts.size(); ++i) {
            auto& in = p.insts[i];
            for (size_t c = 0; c < C; ++c) {
                bool ok = false;
                bool is_pair = c == sc.pair_hi_class;
                if (in.op == OP_PAIRHI) ok = is_pair;
                else if (in.op == OP_ANY) ok = !is_pair;  // only ever positioned on the low half of a pair
                else if (is_pair || rep[c] < 0) ok = false;
                else if (in.op == OP_CHAR) ok = static_cast<uint32_t>(in.a) == static_cast<uint32_t>(rep[c]);
                else if (in.op == OP_SET) {
                    auto& s = p.sets[in.a];
                    uint32_t u = static_cast<uint32_t>(rep[c]);
                    auto it = std::upper_bound(s.begin(), s.end(), std::make_pair(u, 0xFFFFFFFFu));
                    ok = it != s.begin() && (it - 1)->second >= u;
                }
                accepts_[i][c] = ok;
            }
        }
    }

    Tdfa run() {
        const size_t C = sc_.n_classes;
        Item init{0, std::vector<int16_t>(n_slots_, -1)};
        states_.push_back({init});
        index_[key_of(states_[0])].push_back(0);
        out_.op_off = {0, 0};  // list 0 = empty
        oplists_[{}] = 0;
        for (size_t s = 0; s < states_.size(); ++s) {
            for (size_t c = 0; c < C; ++c) step(s, c);
            finals(s);
        }
        out_.n_states = static_cast<uint32_t>(states_.size());
        out_.n_classes = static_cast<uint32_t>(C);
        out_.n_regs = static_cast<uint32_t>(n_regs_);
        out_.n_slots = static_cast<uint32_t>(n_slots_);
        return std::move(out_);
    }

  private:
    const CaptureProgram& p_;
    const SymbolClasses& sc_;
    size_t max_states_, max_regs_, n_slots_;
    std::vector<std::vector<uint8_t>> accepts_;
    std::map<int, std::vector<ClosureEntry>> closures_;
    std::vector<std::vector<Item>> states_;
    std::map<std::string, std::vector<uint32_t>> index_;
    std::map<std::vector<uint16_t>, uint32_t> oplists_;
    size_t n_regs_ = 0;
    Tdfa out_;

    const std::vector<ClosureEntry>& closure(int pc) {
        auto it = closures_.find(pc);
        if (it != closures_.end()) return it->second;
        std::vector<ClosureEntry> out;
        std::vector<uint8_t> seen(p_.insts.size(), 0);
        std::vector<std::pair<int, uint64_t>> stack{{pc, 0}};
        while (!stack.empty()) {
            auto [q, mask] = stack.back();
            stack.pop_back();
            if (seen[q]) continue;
            seen[q] = 1;
            auto& in = p_.insts[q];
            switch (in.op) {
                case OP_JMP: stack.push_back({in.a, mask}); break;
                case OP_SPLIT:
                    stack.push_back({in.b, mask});
                    stack.push_back({in.a, mask});  // preferred branch is explored first
                    break;
                case OP_SAVE: stack.push_back({q + 1, mask | (1ull << in.a)}); break;
                case OP_MATCH: out.push_back({-1, mask}); break;
                default: out.push_back({q, mask}); break;
            }
        }
        return closures_.emplace(pc, std::move(out)).first->second;
    }

    static std::string key_of(const std::vector<Item>& items) {
        std::string k;
        for (auto& it : items) {
            k.append(reinterpret_cast<const char*>(&it.pc), sizeof(int));
            for (int16_t r : it.reg) k.push_back(r < 0 ? '0' : '1');
        }
        return k;
    }

    static constexpr int16_t POS = -2;

    // sequentialise a parallel move set {dst <- src}; src == POS reads the current position
    std::vector<uint16_t> schedule(std::vector<std::pair<int16_t, int16_t>> moves) {
        std::vector<uint16_t> out;
        auto enc = [](int16_t d, int16_t s) { return static_cast<uint16_t>((d << 8) | (s == POS ? 0xFF : s)); };
        while (!moves.empty()) {
            bool progress = false;
            for (size_t i = 0; i < moves.size(); ++i) {
                bool dst_is_read = false;
                for (size_t j = 0; j < moves.size(); ++j)
                    if (j != i && moves[j].second == moves[i].first) { dst_is_read = true; break; }
                if (!dst_is_read) {
                    out.push_back(enc(moves[i].first, moves[i].second));
                    moves.erase(moves.begin() + static_cast<long>(i));
                    progress = true;
                    break;
                }
            }
            if (progress) continue;
            // only cycles remain: park one source in the scratch register
            int16_t tmp = static_cast<int16_t>(max_regs_);  // reserved id, see run-time register file size
            int16_t victim = moves[0].second;
            out.push_back(enc(tmp, victim));
            for (auto& m : moves)
                if (m.second == victim) m.second = tmp;
            uses_tmp_ = true;
        }
        return out;
    }
    bool uses_tmp_ = false;

    uint32_t oplist_id(const std::vector<uint16_t>& ops) {
        auto it = oplists_.find(ops);
        if (it != oplists_.end()) return it->second;
        uint32_t id = static_cast<uint32_t>(oplists_.size());
        if (id >= 0xFFFF) throw UnsupportedError("capture automaton needs too many distinct register programs");
        oplists_.emplace(ops, id);
        out_.ops.insert(out_.ops.end(), ops.begin(), ops.end());
        out_.op_off.push_back(static_cast<uint32_t>(out_.ops.size()));
        return id;
    }

    void note_reg(int16_t r) {
        if (r >= 0 && static_cast<size_t>(r) + 1 > n_regs_) n_regs_ = static_cast<size_t>(r) + 1;
    }

    void step(size_t s, size_t c) {
        // copy: states_ may grow (and reallocate) below
        const std::vector<Item> cur = states_[s];
        struct NewItem { int pc; std::vector<int16_t> src; };  // src per slot: register of `cur`, POS, or -1
        std::vector<NewItem> next;
        std::vector<uint8_t> seen(p_.insts.size(), 0);
        for (auto& it : cur)
            for (auto& ce : closure(it.pc)) {
                if (ce.target < 0 || seen[ce.target]) continue;
                seen[ce.target] = 1;
                if (!accepts_[ce.target][c]) continue;
                NewItem ni{ce.target + 1, it.reg};
                for (size_t k = 0; k < n_slots_; ++k)
                    if (ce.mask >> k & 1) ni.src[k] = POS;
                next.push_back(std::move(ni));
            }
        if (next.empty()) {
            out_.trans.push_back(0xFFFFu);
            return;
        }
        std::vector<Item> shape;
        for (auto& ni : next) {
            Item it{ni.pc, std::vector<int16_t>(n_slots_, -1)};
            for (size_t k = 0; k < n_slots_; ++k) it.reg[k] = ni.src[k] == -1 ? -1 : 0;
            shape.push_back(std::move(it));
        }
        const std::string key = key_of(shape);
        auto& cands = index_[key];
        long best = -1;
        std::vector<std::pair<int16_t, int16_t>> best_moves;
        for (uint32_t cand : cands) {
            const auto& B = states_[cand];
            std::map<int16_t, int16_t> src_of;
            bool ok = true;
            for (size_t j = 0; j < B.size() && ok; ++j)
                for (size_t k = 0; k < n_slots_; ++k) {
                    int16_t b = B[j].reg[k];
                    if (b < 0) continue;
                    int16_t v = next[j].src[k];
                    auto ins = src_of.emplace(b, v);
                    if (!ins.second && ins.first->second != v) { ok = false; break; }
                }
            if (!ok) continue;
            std::vector<std::pair<int16_t, int16_t>> moves;
            for (auto& kv : src_of)
                if (kv.first != kv.second) moves.push_back({kv.first, kv.second});
            if (best < 0 || moves.size() < best_moves.size()) {
                best = cand;
                best_moves = std::move(moves);
            }
        }
        if (best < 0) {
            // new state: inherited registers keep their ids, every slot set right now shares one fresh id
            std::set<int16_t> used;
            for (auto& ni : next)
                for (int16_t v : ni.src)
                    if (v >= 0) used.insert(v);
            int16_t fresh = 0;
            while (used.count(fresh)) ++fresh;
            bool any_pos = false;
            std::vector<Item> items;
            for (auto& ni : next) {
                Item it{ni.pc, std::vector<int16_t>(n_slots_, -1)};
                for (size_t k = 0; k < n_slots_; ++k) {
                    int16_t v = ni.src[k];
                    if (v == POS) { it.reg[k] = fresh; any_pos = true; }
                    else it.reg[k] = v;
                    note_reg(it.reg[k]);
                }
                items.push_back(std::move(it));
            }
            if (n_regs_ > max_regs_) throw UnsupportedError("capture automaton needs too many tag registers");
            if (states_.size() >= max_states_) throw UnsupportedError("capture automaton exceeds the state limit");
            best = static_cast<long>(states_.size());
            states_.push_back(std::move(items));
            cands.push_back(static_cast<uint32_t>(best));
            best_moves.clear();
            if (any_pos) best_moves.push_back({fresh, POS});
        }
        uint32_t ops = oplist_id(schedule(best_moves));
        out_.trans.push_back(static_cast<uint32_t>(best) | (ops << 16));
    }

    void finals(size_t s) {
        const std::vector<Item>& cur = states_[s];
        std::vector<uint8_t> fin(n_slots_, 0xFF);
        bool acc = false;
        for (auto& it : cur) {
            for (auto& ce : closure(it.pc)) {
                if (ce.target >= 0) continue;
                acc = true;
                for (size_t k = 0; k < n_slots_; ++k) {
                    if (ce.mask >> k & 1) fin[k] = 0xFE;
                    else if (it.reg[k] >= 0) fin[k] = static_cast<uint8_t>(it.reg[k]);
                }
                break;
            }
            if (acc) break;
        }
        out_.accepting.push_back(acc ? 1 : 0);
        out_.fin.insert(out_.fin.end(), fin.begin(), fin.end());
    }

  public:
    bool uses_tmp() const { return uses_tmp_; }

    PikeTables tables() {
        PikeTables t;
        t.n_insts = static_cast<uint32_t>(p_.insts.size());
        t.n_slots = static_cast<uint32_t>(n_slots_);
        t.n_classes = sc_.n_classes;
        t.clo_off.push_back(0);
        for (size_t pc = 0; pc < p_.insts.size(); ++pc) {
            for (auto& ce : closure(static_cast<int>(pc))) {
                t.clo_target.push_back(ce.target);
                t.clo_mask.push_back(ce.mask);
            }
            t.clo_off.push_back(static_cast<uint32_t>(t.clo_target.size()));
        }
        for (auto& row : accepts_) t.accepts.insert(t.accepts.end(), row.begin(), row.end());
        return t;
    }
};

}  // namespace

// Moore minimisation of the tagged automaton: two states are merged when they are indistinguishable for every input
// suffix — same acceptance and final register recipe, and for every class the same command list and equivalent
// successors. The command lists act on one global register file, so equal lists mean equal effects. Determinisation
// with register maps leaves many such duplicates (states reached with different histories but the same future).
void minimise_tdfa(Tdfa& t) {
    const size_t S = t.n_states, C = t.n_classes, G = t.n_slots;
    if (S < 2) return;
    std::vector<uint32_t> block(S);
    {
        std::map<std::vector<uint8_t>, uint32_t> ids;
        for (size_t s = 0; s < S; ++s) {
            std::vector<uint8_t> key(t.fin.begin() + static_cast<long>(s * G), t.fin.begin() + static_cast<long>((s + 1) * G));
            key.push_back(t.accepting[s]);
            block[s] = ids.emplace(std::move(key), static_cast<uint32_t>(ids.size())).first->second;
        }
    }
    size_t n_blocks = 0;
    for (;;) {
        std::map<std::vector<uint32_t>, uint32_t> ids;
        std::vector<uint32_t> next_block(S);
        std::vector<uint32_t> key(2 * C + 1);
        for (size_t s = 0; s < S; ++s) {
            key[0] = block[s];
            for (size_t c = 0; c < C; ++c) {
                const uint32_t ent = t.trans[s * C + c], nx = ent & 0xFFFFu;
                key[1 + 2 * c] = nx == 0xFFFFu ? 0xFFFFFFFFu : block[nx];
                key[2 + 2 * c] = nx == 0xFFFFu ? 0u : ent >> 16;
            }
            next_block[s] = ids.emplace(key, static_cast<uint32_t>(ids.size())).first->second;
        }
        const bool stable = ids.size() == n_blocks;
        n_blocks = ids.size();
        block.swap(next_block);
        if (stable) break;
    }
    if (n_blocks == S) return;
    // new ids in order of first occurrence (state 0 stays the start state), one representative per block
    std::vector<uint32_t> id_of_block(n_blocks, 0xFFFFFFFFu), rep;
    for (size_t s = 0; s < S; ++s)
        if (id_of_block[block[s]] == 0xFFFFFFFFu) {
            id_of_block[block[s]] = static_cast<uint32_t>(rep.size());
            rep.push_back(static_cast<uint32_t>(s));
        }
    Tdfa m;
    m.n_states = static_cast<uint32_t>(rep.size());
    m.n_classes = t.n_classes;
    m.n_regs = t.n_regs;
    m.n_slots = t.n_slots;
    m.op_off = t.op_off;
    m.ops = t.ops;
    for (uint32_t s : rep) {
        for (size_t c = 0; c < C; ++c) {
            const uint32_t ent = t.trans[static_cast<size_t>(s) * C + c], nx = ent & 0xFFFFu;
            m.trans.push_back(nx == 0xFFFFu ? ent : (id_of_block[block[nx]] | (ent & 0xFFFF0000u)));
        }
        m.accepting.push_back(t.accepting[s]);
        m.fin.insert(m.fin.end(), t.fin.begin() + static_cast<long>(s * G), t.fin.begin() + static_cast<long>((s + 1) * G));
    }
    t = std::move(m);
}

PikeTables build_pike_tables(const CaptureProgram& p, const SymbolClasses& sc) {
    TdfaBuilder b(p, sc, 1, 1);
    return b.tables();
}

Tdfa placeholder_tdfa(const CaptureProgram& p, const SymbolClasses& sc) {
    Tdfa t;
    t.n_states = 1;
    t.n_classes = sc.n_classes;
    t.n_regs = 0;
    t.n_slots = static_cast<uint32_t>(2 * p.n_groups);
    t.trans.assign(sc.n_classes, 0xFFFFu);
    t.op_off = {0, 0};
    t.accepting = {0};
    t.fin.assign(t.n_slots, 0xFF);
    return t;
}

Tdfa build_tdfa(const CaptureProgram& p, const SymbolClasses& sc, size_t max_states, size_t max_regs) {
    if (max_states > 0xFFFE) max_states = 0xFFFE;
    if (max_regs > 250) max_regs = 250;
    TdfaBuilder b(p, sc, max_states, max_regs);
    Tdfa t = b.run();
    if (b.uses_tmp()) {  // the scratch register was encoded as id `max_regs`; give it the first free id instead
        const uint16_t from = static_cast<uint16_t>(max_regs), to = static_cast<uint16_t>(t.n_regs);
        for (auto& op : t.ops) {
            uint16_t d = op >> 8, s = op & 0xFF;
            if (d == from) d = to;
            if (s == from) s = to;
            op = static_cast<uint16_t>((d << 8) | s);
        }
        t.n_regs += 1;
    }
    return t;
}

}  // namespace gorp
