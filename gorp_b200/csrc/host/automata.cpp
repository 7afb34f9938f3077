#include "automata.hpp"

#include <algorithm>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <set>
#include <unordered_map>

namespace gorp {
namespace {

constexpr uint32_t MAXC = 0xFFFF;
using Ranges = std::vector<std::pair<uint32_t, uint32_t>>;

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
    if (prev <= MAXC) out.push_back({prev, MAXC});
    return out;
}

// ------------------------------------------------------------------ brics RegExp (flags = NONE) -> AST
struct Node;
using NodeP = std::unique_ptr<Node>;
struct Node {
    enum T { Str, Set, Union, Concat, Repeat } t;
    ustring str;
    Ranges set;
    NodeP a, b;
    int min = 0, max = -1;
};

class BricsParser {
  public:
    explicit BricsParser(const ustring& s) : b_(s) {}
    NodeP parse() {
        if (b_.empty()) return str(u"");
        NodeP e = parse_union();
        if (pos_ < b_.size()) throw std::invalid_argument(strfmt("end-of-string expected at position %zu", pos_));
        return e;
    }

  private:
    const ustring& b_;
    size_t pos_ = 0;

    static NodeP str(ustring s) {
        auto n = std::make_unique<Node>();
        n->t = Node::Str;
        n->str = std::move(s);
        return n;
    }
    static NodeP set(Ranges r) {
        auto n = std::make_unique<Node>();
        n->t = Node::Set;
        n->set = normalize(std::move(r));
        return n;
    }
    static NodeP bin(Node::T t, NodeP a, NodeP b) {
        auto n = std::make_unique<Node>();
        n->t = t;
        n->a = std::move(a);
        n->b = std::move(b);
        return n;
    }
    bool more() const { return pos_ < b_.size(); }
    bool peek(const char* chars) const {
        if (!more() || b_[pos_] >= 0x80) return false;
        return std::strchr(chars, static_cast<char>(b_[pos_])) != nullptr && b_[pos_] != 0;
    }
    bool match(char16_t c) {
        if (more() && b_[pos_] == c) { ++pos_; return true; }
        return false;
    }
    char16_t next() {
        if (!more()) throw std::invalid_argument("unexpected end-of-string");
        return b_[pos_++];
    }
    char16_t char_exp() {
        match(u'\\');
        return next();
    }

    NodeP parse_union() {
        NodeP e = parse_concat();
        if (match(u'|')) e = bin(Node::Union, std::move(e), parse_union());
        return e;
    }
    NodeP parse_concat() {
        NodeP e = parse_repeat();
        if (more() && !peek(")|")) e = bin(Node::Concat, std::move(e), parse_concat());
        return e;
    }
    static NodeP rep(NodeP e, int mn, int mx) {
        auto n = std::make_unique<Node>();
        n->t = Node::Repeat;
        n->a = std::move(e);
        n->min = mn;
        n->max = mx;
        return n;
    }
    int number(size_t st) {
        if (pos_ - st > 6) throw std::invalid_argument("repeat count too large");
        int v = 0;
        for (size_t i = st; i < pos_; ++i) v = v * 10 + (b_[i] - u'0');
        return v;
    }
    NodeP parse_repeat() {
        NodeP e = parse_charclass_exp();
        while (peek("?*+{")) {
            if (match(u'?')) e = rep(std::move(e), 0, 1);
            else if (match(u'*')) e = rep(std::move(e), 0, -1);
            else if (match(u'+')) e = rep(std::move(e), 1, -1);
            else if (match(u'{')) {
                size_t st = pos_;
                while (peek("0123456789")) next();
                if (st == pos_) throw std::invalid_argument(strfmt("integer expected at position %zu", pos_));
                int n = number(st), m = -1;
                if (match(u',')) {
                    st = pos_;
                    while (peek("0123456789")) next();
                    if (st != pos_) m = number(st);
                } else {
                    m = n;
                }
                if (!match(u'}')) throw std::invalid_argument(strfmt("expected '}' at position %zu", pos_));
                e = rep(std::move(e), n, m);
            }
        }
        return e;
    }
    NodeP parse_charclass_exp() {
        if (match(u'[')) {
            bool neg = match(u'^');
            Ranges r = parse_charclass();
            while (more() && !peek("]")) {
                Ranges x = parse_charclass();
                r.insert(r.end(), x.begin(), x.end());
            }
            if (neg) r = complement(r);
            if (!match(u']')) throw std::invalid_argument(strfmt("expected ']' at position %zu", pos_));
            return set(std::move(r));
        }
        return parse_simple();
    }
    Ranges parse_charclass() {
        char16_t c = char_exp();
        if (match(u'-')) {
            if (peek("]")) return {{c, c}, {u'-', u'-'}};
            char16_t d = char_exp();
            return {{c, d}};  // reversed range: empty
        }
        return {{c, c}};
    }
    NodeP parse_simple() {
        if (match(u'.')) return set({{0, MAXC}});
        if (match(u'"')) {
            size_t st = pos_;
            while (more() && b_[pos_] != u'"') ++pos_;
            if (!match(u'"')) throw std::invalid_argument(strfmt("expected '\"' at position %zu", pos_));
            return str(b_.substr(st, pos_ - 1 - st));
        }
        if (match(u'(')) {
            if (match(u')')) return str(u"");
            NodeP e = parse_union();
            if (!match(u')')) throw std::invalid_argument(strfmt("expected ')' at position %zu", pos_));
            return e;
        }
        char16_t c = char_exp();
        return set({{c, c}});
    }
};

// ------------------------------------------------------------------ AST -> epsilon-NFA
struct Nfa {
    std::vector<std::vector<int>> eps;
    std::vector<std::vector<std::pair<Ranges, int>>> edges;
    int add() {
        eps.emplace_back();
        edges.emplace_back();
        if (eps.size() > 400000) throw std::invalid_argument("automaton too large (repeat unrolling)");
        return static_cast<int>(eps.size()) - 1;
    }
};

std::pair<int, int> build(Nfa& n, const Node& e) {
    switch (e.t) {
        case Node::Str: {
            int s = n.add(), cur = s;
            for (char16_t ch : e.str) {
                int t = n.add();
                n.edges[cur].push_back({Ranges{{ch, ch}}, t});
                cur = t;
            }
            return {s, cur};
        }
        case Node::Set: {
            int s = n.add(), t = n.add();
            if (!e.set.empty()) n.edges[s].push_back({e.set, t});
            return {s, t};
        }
        case Node::Union: {
            int s = n.add(), t = n.add();
            for (const Node* sub : {e.a.get(), e.b.get()}) {
                auto f = build(n, *sub);
                n.eps[s].push_back(f.first);
                n.eps[f.second].push_back(t);
            }
            return {s, t};
        }
        case Node::Concat: {
            auto f = build(n, *e.a);
            auto g = build(n, *e.b);
            n.eps[f.second].push_back(g.first);
            return {f.first, g.second};
        }
        case Node::Repeat: {
            int s = n.add(), cur = s;
            if (e.max != -1 && e.min > e.max) return {s, n.add()};  // Automaton.repeat(min>max) == empty language
            for (int i = 0; i < e.min; ++i) {
                auto f = build(n, *e.a);
                n.eps[cur].push_back(f.first);
                cur = f.second;
            }
            if (e.max == -1) {
                auto f = build(n, *e.a);
                int loop = n.add();
                n.eps[cur].push_back(loop);
                n.eps[loop].push_back(f.first);
                n.eps[f.second].push_back(loop);
                return {s, loop};
            }
            int end = n.add();
            n.eps[cur].push_back(end);
            for (int i = e.min; i < e.max; ++i) {
                auto f = build(n, *e.a);
                n.eps[cur].push_back(f.first);
                cur = f.second;
                n.eps[cur].push_back(end);
            }
            return {s, end};
        }
    }
    return {0, 0};
}

}  // namespace

int MinDfa::step(int s, uint32_t c) const {
    const auto& row = trans[s];
    size_t lo = 0, hi = row.size();
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (row[mid].hi < c) lo = mid + 1;
        else hi = mid;
    }
    if (lo < row.size() && row[lo].lo <= c) return static_cast<int>(row[lo].to);
    return -1;
}

std::vector<uint32_t> MinDfa::start_points() const {
    std::set<uint32_t> pts{0};
    for (auto& row : trans)
        for (auto& t : row) {
            pts.insert(t.lo);
            if (t.hi < MAXC) pts.insert(t.hi + 1);
        }
    return {pts.begin(), pts.end()};
}

MinDfa brics_min_dfa(const ustring& regex) {
    NodeP ast = BricsParser(regex).parse();
    Nfa nfa;
    auto frag = build(nfa, *ast);
    const int start = frag.first, final_state = frag.second;

    // alphabet partition local to this regex
    std::set<uint32_t> cutset{0};
    for (auto& es : nfa.edges)
        for (auto& e : es)
            for (auto& r : e.first) {
                cutset.insert(r.first);
                if (r.second < MAXC) cutset.insert(r.second + 1);
            }
    std::vector<uint32_t> cuts(cutset.begin(), cutset.end());
    const int ncls = static_cast<int>(cuts.size());
    // per NFA state: (class -> destinations)
    std::vector<std::vector<std::pair<int, int>>> moves(nfa.eps.size());
    for (size_t s = 0; s < nfa.edges.size(); ++s)
        for (auto& e : nfa.edges[s])
            for (auto& r : e.first) {
                int a = static_cast<int>(std::lower_bound(cuts.begin(), cuts.end(), r.first) - cuts.begin());
                int b = static_cast<int>(std::upper_bound(cuts.begin(), cuts.end(), r.second) - cuts.begin());
                for (int c = a; c < b; ++c) moves[s].push_back({c, e.second});
            }
    auto closure = [&](std::vector<int> seed) {
        std::vector<char> seen(nfa.eps.size(), 0);
        std::vector<int> stack;
        for (int s : seed)
            if (!seen[s]) { seen[s] = 1; stack.push_back(s); }
        std::vector<int> out;
        while (!stack.empty()) {
            int s = stack.back();
            stack.pop_back();
            out.push_back(s);
            for (int t : nfa.eps[s])
                if (!seen[t]) { seen[t] = 1; stack.push_back(t); }
        }
        std::sort(out.begin(), out.end());
        return out;
    };
    std::map<std::vector<int>, int> ids;
    std::vector<std::vector<int>> order;
    std::vector<std::vector<int>> table;
    ids[closure({start})] = 0;
    order.push_back(closure({start}));
    for (size_t i = 0; i < order.size(); ++i) {
        if (order.size() > 200000) throw std::invalid_argument("automaton too large");
        std::vector<std::vector<int>> dst(ncls);
        for (int s : order[i])
            for (auto& m : moves[s]) dst[m.first].push_back(m.second);
        std::vector<int> row(ncls, -1);
        for (int c = 0; c < ncls; ++c) {
            if (dst[c].empty()) continue;
            auto key = closure(dst[c]);
            auto it = ids.find(key);
            if (it == ids.end()) {
                it = ids.emplace(key, static_cast<int>(order.size())).first;
                order.push_back(key);
            }
            row[c] = it->second;
        }
        table.push_back(std::move(row));
    }
    const int n = static_cast<int>(order.size());
    std::vector<char> acc(n + 1, 0);
    for (int s = 0; s < n; ++s) acc[s] = std::binary_search(order[s].begin(), order[s].end(), final_state);
    // Moore refinement on the total DFA (state n == dead)
    table.push_back(std::vector<int>(ncls, n));
    for (auto& row : table)
        for (int& t : row)
            if (t < 0) t = n;
    std::vector<int> part(n + 1);
    for (int s = 0; s <= n; ++s) part[s] = acc[s] ? 1 : 0;
    size_t nblocks = 0;
    {
        std::set<int> u(part.begin(), part.end());
        nblocks = u.size();
    }
    for (;;) {
        std::map<std::vector<int>, int> sig;
        std::vector<int> np(n + 1);
        std::vector<int> key(ncls + 1);
        for (int s = 0; s <= n; ++s) {
            key[0] = part[s];
            for (int c = 0; c < ncls; ++c) key[c + 1] = part[table[s][c]];
            np[s] = sig.emplace(key, static_cast<int>(sig.size())).first->second;
        }
        part.swap(np);
        if (sig.size() == nblocks) break;
        nblocks = sig.size();
    }
    std::vector<int> rep(nblocks, -1);
    for (int s = 0; s <= n; ++s)
        if (rep[part[s]] < 0) rep[part[s]] = s;
    auto bt = [&](int b, int c) { return part[table[rep[b]][c]]; };
    std::vector<char> live(nblocks, 0);
    for (size_t b = 0; b < nblocks; ++b) live[b] = acc[rep[b]];
    for (bool changed = true; changed;) {
        changed = false;
        for (size_t b = 0; b < nblocks; ++b)
            if (!live[b])
                for (int c = 0; c < ncls; ++c)
                    if (live[bt(static_cast<int>(b), c)]) { live[b] = 1; changed = true; break; }
    }
    std::vector<int> num(nblocks, -1), olist;
    num[part[0]] = 0;
    olist.push_back(part[0]);
    for (size_t i = 0; i < olist.size(); ++i)
        for (int c = 0; c < ncls; ++c) {
            int t = bt(olist[i], c);
            if (live[t] && num[t] < 0) {
                num[t] = static_cast<int>(olist.size());
                olist.push_back(t);
            }
        }
    MinDfa out;
    for (int b : olist) {
        std::vector<Interval> row;
        for (int c = 0; c < ncls;) {
            int t = bt(b, c);
            if (!live[t]) { ++c; continue; }
            int c2 = c;
            while (c2 + 1 < ncls && bt(b, c2 + 1) == t) ++c2;
            uint32_t hi = (c2 + 1 < ncls) ? cuts[c2 + 1] - 1 : MAXC;
            row.push_back({cuts[c], hi, static_cast<uint32_t>(num[t])});
            c = c2 + 1;
        }
        out.trans.push_back(std::move(row));
        out.accept.push_back(acc[rep[b]] ? 1 : 0);
    }
    return out;
}

// ------------------------------------------------------------------ product (Automata.construct)
namespace {
struct KeyHash {
    const std::vector<int32_t>* arena;
    size_t n;
    size_t operator()(uint32_t id) const {
        const int32_t* p = arena->data() + static_cast<size_t>(id) * n;
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < n; ++i) {
            h ^= static_cast<uint32_t>(p[i]);
            h *= 1099511628211ull;
        }
        return static_cast<size_t>(h ^ (h >> 29));
    }
};
struct KeyEq {
    const std::vector<int32_t>* arena;
    size_t n;
    bool operator()(uint32_t a, uint32_t b) const {
        return std::memcmp(arena->data() + static_cast<size_t>(a) * n, arena->data() + static_cast<size_t>(b) * n,
                           n * sizeof(int32_t)) == 0;
    }
};
}  // namespace

DfaTables build_product(const std::vector<MinDfa>& dfas, size_t max_states) {
    const size_t N = dfas.size();
    std::set<uint32_t> ptset;
    for (auto& d : dfas)
        for (uint32_t p : d.start_points()) ptset.insert(p);  // pointsUnion
    std::vector<uint32_t> points(ptset.begin(), ptset.end());
    const size_t P = points.size();
    // dense per-component tables over the union points
    std::vector<std::vector<int32_t>> tbl(N);
    for (size_t i = 0; i < N; ++i) {
        size_t ns = dfas[i].trans.size();
        tbl[i].assign(ns * P, -1);
        for (size_t s = 0; s < ns; ++s)
            for (size_t c = 0; c < P; ++c) tbl[i][s * P + c] = dfas[i].step(static_cast<int>(s), points[c]);
    }
    std::vector<int32_t> arena(N, 0);  // state 0 = all components initial
    KeyHash kh{&arena, N};
    KeyEq ke{&arena, N};
    std::unordered_map<uint32_t, char, KeyHash, KeyEq> index(1024, kh, ke);
    index.emplace(0u, 0);
    DfaTables out;
    out.n_regex = static_cast<uint32_t>(N);
    out.n_classes = static_cast<uint32_t>(P);
    size_t n_states = 1;
    std::vector<int32_t> cand(N);
    for (size_t v = 0; v < n_states; ++v) {  // ids are assigned in BFS discovery order == visiting order
        for (size_t c = 0; c < P; ++c) {
            bool all_null = true;
            const int32_t* cur = arena.data() + v * N;
            for (size_t i = 0; i < N; ++i) {
                int32_t s = cur[i];
                int32_t t = s < 0 ? -1 : tbl[i][static_cast<size_t>(s) * P + c];
                cand[i] = t;
                all_null &= (t < 0);
            }
            if (all_null) { out.trans.push_back(-1); continue; }
            // tentatively append, look up, roll back when known
            arena.insert(arena.end(), cand.begin(), cand.end());
            uint32_t id = static_cast<uint32_t>(n_states);
            auto it = index.find(id);
            if (it == index.end()) {
                if (n_states >= max_states) throw UnsupportedError("combined DFA exceeds the state limit");
                index.emplace(id, 0);
                ++n_states;
                out.trans.push_back(static_cast<int32_t>(id));
            } else {
                arena.resize(n_states * N);
                out.trans.push_back(static_cast<int32_t>(it->first));
            }
        }
    }
    out.n_states = static_cast<uint32_t>(n_states);
    out.accept_first.assign(n_states, -1);
    out.accept_off.assign(n_states + 1, 0);
    for (size_t v = 0; v < n_states; ++v) {
        const int32_t* cur = arena.data() + v * N;
        for (size_t i = 0; i < N; ++i)
            if (cur[i] >= 0 && dfas[i].accept[cur[i]]) {
                if (out.accept_first[v] < 0) out.accept_first[v] = static_cast<int32_t>(i);
                out.accept_list.push_back(static_cast<int32_t>(i));
            }
        out.accept_off[v + 1] = static_cast<uint32_t>(out.accept_list.size());
    }
    // alphabet(points): char -> index of the greatest point <= char
    out.classmap.assign(65536, 0);
    size_t k = 0;
    for (uint32_t ch = 0; ch < 65536; ++ch) {
        if (k + 1 < P && ch == points[k + 1]) ++k;
        out.classmap[ch] = static_cast<uint16_t>(k);
    }
    return out;
}

CompactDfa compact_tables(const DfaTables& t) {
    const size_t S = t.n_states, C = t.n_classes;
    // 1. merge identical columns
    std::map<std::vector<int32_t>, uint32_t> colid;
    std::vector<uint32_t> colmap(C);
    std::vector<size_t> colrep;
    std::vector<int32_t> col(S);
    for (size_t c = 0; c < C; ++c) {
        for (size_t s = 0; s < S; ++s) col[s] = t.trans[s * C + c];
        auto it = colid.find(col);
        if (it == colid.end()) {
            it = colid.emplace(col, static_cast<uint32_t>(colrep.size())).first;
            colrep.push_back(c);
        }
        colmap[c] = it->second;
    }
    const size_t C2 = colrep.size();
    // 2. Moore refinement seeded by accept_first; block S == dead
    std::vector<int32_t> part(S + 1);
    {
        std::map<int32_t, int32_t> seed;
        seed[-1] = 0;  // dead shares the initial block of the non-accepting states
        for (size_t s = 0; s < S; ++s) part[s] = seed.emplace(t.accept_first[s], static_cast<int32_t>(seed.size())).first->second;
        part[S] = 0;
    }
    size_t nblocks = 0;
    {
        std::set<int32_t> u(part.begin(), part.end());
        nblocks = u.size();
    }
    for (;;) {
        std::unordered_map<std::string, int32_t> sig;
        std::vector<int32_t> np(S + 1);
        std::vector<int32_t> key(C2 + 1);
        for (size_t s = 0; s <= S; ++s) {
            key[0] = part[s];
            for (size_t c = 0; c < C2; ++c) {
                int32_t d = s == S ? -1 : t.trans[s * C + colrep[c]];
                key[c + 1] = part[d < 0 ? S : static_cast<size_t>(d)];
            }
            std::string k(reinterpret_cast<const char*>(key.data()), key.size() * sizeof(int32_t));
            np[s] = sig.emplace(std::move(k), static_cast<int32_t>(sig.size())).first->second;
        }
        part.swap(np);
        if (sig.size() == nblocks) break;
        nblocks = sig.size();
    }
    const int32_t dead = part[S];
    // the dead block may also hold live-numbered states that can never accept: they become -1 too
    std::vector<int32_t> num(nblocks, -1);
    std::vector<size_t> reps;
    CompactDfa out;
    if (part[0] == dead) {  // empty language overall: single non-accepting start state
        out.n_states = 1;
        out.n_classes = static_cast<uint32_t>(C2);
        out.trans.assign(C2, -1);
        out.accept_first.assign(1, -1);
    } else {
        num[part[0]] = 0;
        reps.push_back(0);
        for (size_t i = 0; i < reps.size(); ++i)
            for (size_t c = 0; c < C2; ++c) {
                int32_t d = t.trans[reps[i] * C + colrep[c]];
                if (d < 0) continue;
                int32_t b = part[d];
                if (b == dead || num[b] >= 0) continue;
                num[b] = static_cast<int32_t>(reps.size());
                reps.push_back(static_cast<size_t>(d));
            }
        out.n_states = static_cast<uint32_t>(reps.size());
        out.n_classes = static_cast<uint32_t>(C2);
        out.trans.assign(reps.size() * C2, -1);
        out.accept_first.resize(reps.size());
        for (size_t i = 0; i < reps.size(); ++i) {
            out.accept_first[i] = t.accept_first[reps[i]];
            for (size_t c = 0; c < C2; ++c) {
                int32_t d = t.trans[reps[i] * C + colrep[c]];
                if (d >= 0 && part[d] != dead) out.trans[i * C2 + c] = num[part[d]];
            }
        }
    }
    out.classmap.resize(65536);
    for (size_t ch = 0; ch < 65536; ++ch) out.classmap[ch] = static_cast<uint16_t>(colmap[t.classmap[ch]]);
    return out;
}

}  // namespace gorp
