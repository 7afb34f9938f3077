#include "tails.hpp"

#include <algorithm>
#include <map>
#include <set>

namespace gorp {

namespace {

constexpr uint32_t kDeadT = 0xFFFFu;  // the capture automaton is dead (CAPTURE_FAIL if the DFA side accepts)

struct Refuse {
    std::string why;
};

// minimal DFA of the language of extraction e, recovered from the product tables: acceptance "e in _accept[state]",
// trimmed to the states that can still reach such a state, then minimised (compact_tables)
CompactDfa component_dfa(const DfaTables& R, uint32_t e, const std::vector<uint32_t>& radj_off, const std::vector<uint32_t>& radj) {
    const size_t S = R.n_states, C = R.n_classes;
    std::vector<uint8_t> acc(S, 0), live(S, 0);
    std::vector<uint32_t> stack;
    for (size_t q = 0; q < S; ++q)
        for (uint32_t i = R.accept_off[q]; i < R.accept_off[q + 1]; ++i)
            if (static_cast<uint32_t>(R.accept_list[i]) == e) {
                acc[q] = 1;
                live[q] = 1;
                stack.push_back(static_cast<uint32_t>(q));
            }
    while (!stack.empty()) {  // backwards: everything that can reach an accepting state
        const uint32_t q = stack.back();
        stack.pop_back();
        for (uint32_t i = radj_off[q]; i < radj_off[q + 1]; ++i)
            if (!live[radj[i]]) {
                live[radj[i]] = 1;
                stack.push_back(radj[i]);
            }
    }
    DfaTables sub;
    sub.n_classes = static_cast<uint32_t>(C);
    sub.classmap = R.classmap;
    sub.n_regex = 1;
    std::vector<int32_t> id(S, -1);
    std::vector<uint32_t> order;
    if (live[0]) {  // forwards from the start state, inside the live set
        id[0] = 0;
        order.push_back(0);
        for (size_t i = 0; i < order.size(); ++i)
            for (size_t c = 0; c < C; ++c) {
                const int32_t t = R.trans[static_cast<size_t>(order[i]) * C + c];
                if (t < 0 || !live[t] || id[t] >= 0) continue;
                id[t] = static_cast<int32_t>(order.size());
                order.push_back(static_cast<uint32_t>(t));
            }
    }
    if (order.empty()) {  // empty language: one state, nothing accepted
        sub.n_states = 1;
        sub.trans.assign(C, -1);
        sub.accept_first.assign(1, -1);
        return compact_tables(sub);
    }
    sub.n_states = static_cast<uint32_t>(order.size());
    sub.trans.assign(order.size() * C, -1);
    sub.accept_first.resize(order.size());
    for (size_t i = 0; i < order.size(); ++i) {
        sub.accept_first[i] = acc[order[i]] ? 0 : -1;
        for (size_t c = 0; c < C; ++c) {
            const int32_t t = R.trans[static_cast<size_t>(order[i]) * C + c];
            if (t >= 0 && live[t]) sub.trans[i * C + c] = id[t];
        }
    }
    return compact_tables(sub);
}

TailAutomaton build_tail(const CompactDfa& D, const Tdfa& T, uint32_t e, uint32_t n_cap_classes, const std::vector<uint32_t>& col_rep_unit,
                         const std::vector<uint32_t>& col_cap, uint32_t width, size_t max_states, size_t max_op_slots) {
    TailAutomaton A;
    try {
        for (uint16_t op : T.ops)
            if ((op & 0xFF) != 0xFF) throw Refuse{"capture automaton copies tag registers"};
        const size_t W = col_rep_unit.size();  // used columns (the rest of `width` is padding)
        const uint32_t Cd = D.n_classes;
        // ---- product BFS over (DFA state, capture state)
        std::map<std::pair<uint32_t, uint32_t>, uint32_t> index;
        std::vector<std::pair<uint32_t, uint32_t>> states;
        std::map<uint32_t, uint32_t> slot_of;  // capture op-list id -> op slot (1..)
        std::vector<uint32_t> slot_def{0};
        std::vector<uint32_t> trans;            // raw product [P * W]
        index.emplace(std::make_pair(0u, 0u), 0u);
        states.push_back({0u, 0u});
        for (size_t p = 0; p < states.size(); ++p) {
            const uint32_t d = states[p].first, t = states[p].second;
            for (size_t k = 0; k < W; ++k) {
                if (k == 0x0A) {  // the terminator column: laid out by the engine (outcome rows), never a transition
                    trans.push_back(0xFFFFu);
                    continue;
                }
                const int32_t d2 = D.trans[static_cast<size_t>(d) * Cd + D.classmap[col_rep_unit[k]]];
                if (d2 < 0) {
                    trans.push_back(0xFFFFu);
                    continue;
                }
                uint32_t t2 = kDeadT, slot = 0;
                if (t != kDeadT) {
                    const uint32_t ent = T.trans[static_cast<size_t>(t) * n_cap_classes + col_cap[k]];
                    if ((ent & 0xFFFFu) != 0xFFFFu) {
                        t2 = ent & 0xFFFFu;
                        const uint32_t ol = ent >> 16;
                        if (T.op_off[ol + 1] > T.op_off[ol]) {
                            auto it = slot_of.find(ol);
                            if (it == slot_of.end()) {
                                if (slot_def.size() > max_op_slots) throw Refuse{"too many distinct register command lists"};
                                it = slot_of.emplace(ol, static_cast<uint32_t>(slot_def.size())).first;
                                slot_def.push_back(ol);
                            }
                            slot = it->second;
                        }
                    }
                }
                const auto key = std::make_pair(static_cast<uint32_t>(d2), t2);
                auto it = index.find(key);
                if (it == index.end()) {
                    if (states.size() >= 16 * max_states + 64) throw Refuse{"product automaton exceeds the state limit"};
                    it = index.emplace(key, static_cast<uint32_t>(states.size())).first;
                    states.push_back(key);
                }
                trans.push_back(it->second | (slot << 16));
            }
        }
        const size_t P = states.size();

        // ---- writers of each capture register: r -> op slots
        std::map<uint32_t, std::vector<uint32_t>> writers;
        for (uint32_t s = 1; s < slot_def.size(); ++s) {
            const uint32_t ol = slot_def[s];
            for (uint32_t o = T.op_off[ol]; o < T.op_off[ol + 1]; ++o) {
                auto& w = writers[static_cast<uint32_t>(T.ops[o] >> 8)];
                if (std::find(w.begin(), w.end(), s) == w.end()) w.push_back(s);
            }
        }

        // ---- outcomes
        std::map<std::vector<uint32_t>, uint32_t> outcome_id;
        A.outcomes.push_back({-1, 0});
        std::vector<uint32_t> outcome_of(P, 0);
        for (size_t p = 0; p < P; ++p) {
            if (D.accept_first[states[p].first] < 0) continue;  // the DFA side rejects: MISS
            const uint32_t t = states[p].second;
            std::vector<uint32_t> key;
            if (t == kDeadT || !T.accepting[t]) {
                key = {0xFFFFFFFEu};
            } else {
                key = {0u};
                for (uint32_t k = 0; k < T.n_slots; ++k) key.push_back(T.fin[static_cast<size_t>(t) * T.n_slots + k]);
            }
            auto it = outcome_id.find(key);
            if (it == outcome_id.end()) {
                FusedAutomaton::Outcome o{};
                if (key[0] == 0xFFFFFFFEu) {
                    o.ext_code = -2 - static_cast<int32_t>(e);
                } else {
                    o.ext_code = static_cast<int32_t>(e);
                    o.res_off = static_cast<uint32_t>(A.res.size());
                    for (uint32_t k = 0; k < T.n_slots; ++k) {
                        const uint32_t f = key[1 + k];
                        uint32_t packed = 0;
                        if (f == 0xFE) {
                            packed = FusedAutomaton::kLenSlot;
                        } else if (f != 0xFF) {
                            auto w = writers.find(f);
                            if (w != writers.end()) {
                                if (w->second.size() > 4) throw Refuse{"a group boundary has more than 4 writers"};
                                for (size_t i = 0; i < w->second.size(); ++i) packed |= w->second[i] << (8 * i);
                                if (w->second.size() > 1)
                                    for (uint32_t s : w->second)
                                        if (std::find(A.init_slots.begin(), A.init_slots.end(), s) == A.init_slots.end()) A.init_slots.push_back(s);
                            }
                        }
                        A.res.push_back(packed);
                    }
                }
                if (A.outcomes.size() >= 250) throw Refuse{"too many distinct outcomes"};
                it = outcome_id.emplace(key, static_cast<uint32_t>(A.outcomes.size())).first;
                A.outcomes.push_back(o);
            }
            outcome_of[p] = it->second;
        }

        // ---- Moore minimisation: same outcome and, per column, same op slot and equivalent successor (block P == dead)
        std::vector<uint32_t> part(P + 1);
        for (size_t p = 0; p < P; ++p) part[p] = outcome_of[p] + 1;
        part[P] = 0;
        size_t nblocks = 0;
        {
            std::set<uint32_t> u(part.begin(), part.end());
            nblocks = u.size();
        }
        for (;;) {
            std::map<std::vector<uint32_t>, uint32_t> sig;
            std::vector<uint32_t> np(P + 1), key(W + 1);
            for (size_t p = 0; p <= P; ++p) {
                key[0] = part[p];
                for (size_t k = 0; k < W; ++k) {
                    const uint32_t ent = p == P ? 0xFFFFu : trans[p * W + k];
                    const uint32_t nx = ent & 0xFFFFu;
                    key[k + 1] = nx == 0xFFFFu ? (part[P] | 0x80000000u) : (part[nx] | ((ent >> 16) << 16));
                }
                np[p] = sig.emplace(key, static_cast<uint32_t>(sig.size())).first->second;
            }
            part.swap(np);
            if (sig.size() == nblocks) break;
            nblocks = sig.size();
        }
        const uint32_t dead_block = part[P];
        std::vector<int64_t> num(nblocks, -1);
        std::vector<size_t> rep;
        if (part[0] == dead_block) {  // nothing can ever be accepted: one state, every transition dead
            A.n_states = 1;
            A.trans.assign(width, 0xFFFFu);
            A.outcome_of.assign(1, 0);
        } else {
            num[part[0]] = 0;
            rep.push_back(0);
            for (size_t i = 0; i < rep.size(); ++i)
                for (size_t k = 0; k < W; ++k) {
                    const uint32_t ent = trans[rep[i] * W + k];
                    if ((ent & 0xFFFFu) == 0xFFFFu) continue;
                    const uint32_t b = part[ent & 0xFFFFu];
                    if (b == dead_block || num[b] >= 0) continue;
                    num[b] = static_cast<int64_t>(rep.size());
                    rep.push_back(ent & 0xFFFFu);
                }
            if (rep.size() > max_states) throw Refuse{"tail automaton exceeds the state limit"};
            A.n_states = static_cast<uint32_t>(rep.size());
            A.trans.assign(rep.size() * width, 0xFFFFu);
            A.outcome_of.resize(rep.size());
            for (size_t i = 0; i < rep.size(); ++i) {
                A.outcome_of[i] = outcome_of[rep[i]];
                for (size_t k = 0; k < W; ++k) {
                    const uint32_t ent = trans[rep[i] * W + k];
                    const uint32_t nx = ent & 0xFFFFu;
                    if (nx != 0xFFFFu && part[nx] != dead_block) A.trans[i * width + k] = static_cast<uint32_t>(num[part[nx]]) | (ent & 0xFFFF0000u);
                }
            }
        }
        A.n_op_slots = static_cast<uint32_t>(slot_def.size() - 1);
        A.n_boundaries = T.n_slots;
        A.available = true;
    } catch (const Refuse& r) {
        A = TailAutomaton{};
        A.why_not = r.why;
    }
    return A;
}

}  // namespace

TailSet build_tails(const CompiledDefinition& def, const DeviceModel& m, size_t max_states, size_t max_op_slots) {
    TailSet T;
    const DfaTables& R = def.dfa;
    const size_t S = R.n_states, C = R.n_classes, E = m.tdfas.size();
    if (E == 0 || m.tdfas.size() != m.n_groups.size() || def.extractions.size() != E) return T;
    const uint32_t Cn = m.symbols.n_classes, PAIR = m.symbols.pair_hi_class;

    // ---- columns shared by all tails
    std::map<std::pair<uint32_t, uint32_t>, uint16_t> colid;
    std::vector<uint32_t> col_rep_unit(128), col_cap(128);
    for (uint32_t u = 0; u < 128; ++u) {
        col_rep_unit[u] = u;
        col_cap[u] = m.symbols.classmap[u];
    }
    auto col = [&](uint32_t rep_unit, uint32_t cap_cls) -> uint16_t {
        const auto key = std::make_pair(static_cast<uint32_t>(R.classmap[rep_unit]), cap_cls);
        auto it = colid.find(key);
        if (it != colid.end()) return it->second;
        it = colid.emplace(key, static_cast<uint16_t>(col_rep_unit.size())).first;
        col_rep_unit.push_back(rep_unit);
        col_cap.push_back(cap_cls);
        return it->second;
    };
    T.xcol.resize(65536);
    for (uint32_t u = 0; u < 128; ++u) T.xcol[u] = static_cast<uint16_t>(u);
    for (uint32_t u = 128; u < 65536; ++u) T.xcol[u] = col(u, m.symbols.classmap[u]);
    std::vector<std::pair<uint16_t, uint16_t>> pairs;
    for (uint32_t u = 0xD800; u < 0xDC00; ++u) pairs.push_back({T.xcol[u], col(u, PAIR)});
    T.nl_data_col = col(0x0A, m.symbols.classmap[0x0A]);
    if (col_rep_unit.size() > 1024) return T;
    T.width = static_cast<uint32_t>((col_rep_unit.size() + 3) & ~size_t(3));
    T.pair_col.resize(T.width);
    for (uint32_t k = 0; k < T.width; ++k) T.pair_col[k] = static_cast<uint16_t>(k);
    for (auto& p : pairs) T.pair_col[p.first] = p.second;

    // ---- reverse edges of the product automaton (shared by all component extractions)
    std::vector<uint32_t> radj_off(S + 1, 0), radj;
    for (size_t q = 0; q < S; ++q)
        for (size_t c = 0; c < C; ++c) {
            const int32_t t = R.trans[q * C + c];
            if (t >= 0) ++radj_off[static_cast<size_t>(t) + 1];
        }
    for (size_t q = 0; q < S; ++q) radj_off[q + 1] += radj_off[q];
    radj.resize(radj_off[S]);
    {
        std::vector<uint32_t> fill(radj_off.begin(), radj_off.end() - 1);
        for (size_t q = 0; q < S; ++q)
            for (size_t c = 0; c < C; ++c) {
                const int32_t t = R.trans[q * C + c];
                if (t >= 0) radj[fill[t]++] = static_cast<uint32_t>(q);
            }
    }

    // ---- one tail per extraction
    T.tails.resize(E);
    for (size_t e = 0; e < E; ++e) {
        const CompactDfa De = component_dfa(R, static_cast<uint32_t>(e), radj_off, radj);
        T.tails[e] = build_tail(De, m.tdfas[e], static_cast<uint32_t>(e), Cn, col_rep_unit, col_cap, T.width, max_states, max_op_slots);
        T.any = T.any || T.tails[e].available;
    }

    // ---- where the combined-DFA walk may stop: compact states from which exactly one extraction can still be the first
    //      accepting index, and that extraction has a tail
    const CompactDfa& D = m.dfa;
    const size_t Sc = D.n_states, Cc = D.n_classes, Wd = (E + 63) / 64;
    std::vector<uint64_t> reach(Sc * Wd, 0);
    for (size_t q = 0; q < Sc; ++q)
        if (D.accept_first[q] >= 0) reach[q * Wd + D.accept_first[q] / 64] |= 1ull << (D.accept_first[q] % 64);
    for (bool changed = true; changed;) {
        changed = false;
        for (size_t q = Sc; q-- > 0;)
            for (size_t c = 0; c < Cc; ++c) {
                const int32_t t = D.trans[q * Cc + c];
                if (t < 0) continue;
                for (size_t w = 0; w < Wd; ++w) {
                    const uint64_t v = reach[q * Wd + w] | reach[static_cast<size_t>(t) * Wd + w];
                    if (v != reach[q * Wd + w]) {
                        reach[q * Wd + w] = v;
                        changed = true;
                    }
                }
            }
    }
    T.cut_of_state.assign(Sc, -1);
    for (size_t q = 0; q < Sc; ++q) {
        int count = 0, which = -1;
        for (size_t w = 0; w < Wd; ++w) {
            const uint64_t v = reach[q * Wd + w];
            if (!v) continue;
            count += __builtin_popcountll(v);
            which = static_cast<int>(w * 64) + __builtin_ctzll(v);
        }
        if (count == 1 && T.tails[static_cast<size_t>(which)].available) {
            T.cut_of_state[q] = which;
            ++T.n_cut_states;
        }
    }
    if (T.cut_of_state[0] >= 0) {  // a definition with a single live extraction: keep the start state (row 0 must exist)
        T.cut_of_state[0] = -1;
        --T.n_cut_states;
    }
    return T;
}

TailSet build_fused_tailset(const FusedAutomaton& A, const DeviceModel& m, uint32_t span_stride) {
    TailSet FT;
    if (!A.available) return FT;
    const uint32_t Sx = A.n_states, J = A.n_jcls;
    std::vector<int32_t> col_of_j(J, -1);
    std::vector<uint32_t> j_of_col;
    auto col = [&](uint32_t j) {
        if (col_of_j[j] < 0) {
            col_of_j[j] = static_cast<int32_t>(128 + j_of_col.size());
            j_of_col.push_back(j);
        }
        return static_cast<uint16_t>(col_of_j[j]);
    };
    FT.xcol.resize(65536);
    for (uint32_t u = 0; u < 128; ++u) FT.xcol[u] = static_cast<uint16_t>(u);
    for (uint32_t u = 128; u < 65536; ++u) FT.xcol[u] = col(A.jcls[u]);
    for (uint32_t u = 0xD800; u < 0xDC00; ++u) col(A.pair_of[A.jcls[u]]);
    FT.nl_data_col = col(A.jcls[0x0A]);
    for (size_t k = 0; k < j_of_col.size(); ++k) col(A.pair_of[j_of_col[k]]);  // (pairs of pairs: the identity, no new column)
    FT.width = static_cast<uint32_t>((128 + j_of_col.size() + 3) & ~size_t(3));
    FT.pair_col.resize(FT.width);
    for (uint32_t k = 0; k < FT.width; ++k) FT.pair_col[k] = static_cast<uint16_t>(k);
    for (size_t k = 0; k < j_of_col.size(); ++k) FT.pair_col[128 + k] = col(A.pair_of[j_of_col[k]]);
    TailAutomaton TA;
    TA.available = true;
    TA.n_states = Sx;
    TA.trans.assign(static_cast<size_t>(Sx) * FT.width, 0xFFFFu);
    for (uint32_t r = 0; r < Sx; ++r)
        for (uint32_t k = 0; k < FT.width; ++k) {
            if (k >= 128 && k - 128 >= j_of_col.size()) continue;  // padding column: dead
            const uint32_t j = k < 128 ? A.jcls[k] : j_of_col[k - 128];
            TA.trans[static_cast<size_t>(r) * FT.width + k] = A.trans[static_cast<size_t>(r) * J + j];
        }
    TA.n_op_slots = A.n_op_slots;
    TA.n_boundaries = std::max(span_stride, 1u);
    TA.outcome_of = A.outcome_of;
    for (const FusedAutomaton::Outcome& oc : A.outcomes) {  // recipes padded to the row width: 0 = no writer
        FusedAutomaton::Outcome o2{oc.ext_code, static_cast<uint32_t>(TA.res.size())};
        const uint32_t n = oc.ext_code >= 0 ? 2 * m.n_groups[oc.ext_code] : 0u;
        for (uint32_t k = 0; k < TA.n_boundaries; ++k) {
            const uint32_t packed = k < n ? A.res[oc.res_off + k] : 0u;
            TA.res.push_back(packed);
            if (packed >= 256u)  // several writers: their slots are read through a maximum and must be reset per line
                for (uint32_t sh = 0; sh < 32 && ((packed >> sh) & 0xFFu); sh += 8) {
                    const uint32_t id = (packed >> sh) & 0xFFu;
                    if (id != FusedAutomaton::kLenSlot && std::find(TA.init_slots.begin(), TA.init_slots.end(), id) == TA.init_slots.end())
                        TA.init_slots.push_back(id);
                }
        }
        TA.outcomes.push_back(o2);
    }
    FT.any = true;
    FT.tails.push_back(std::move(TA));
    return FT;
}

}  // namespace gorp
