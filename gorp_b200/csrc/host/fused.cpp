#include "fused.hpp"

#include <algorithm>
#include <map>
#include <set>

namespace gorp {

namespace {

constexpr uint32_t kDeadT = 0xFFFFu;  // capture automaton of this extraction is dead (CAPTURE_FAIL if it wins)

struct Refuse {
    std::string why;
};

}  // namespace

FusedAutomaton build_fused(const DeviceModel& m, size_t max_states, size_t max_op_slots) {
    FusedAutomaton A;
    try {
        const CompactDfa& D = m.dfa;
        const size_t S = D.n_states, C = D.n_classes, E = m.tdfas.size();
        if (E == 0 || m.tdfas.size() != m.n_groups.size()) throw Refuse{"no capture automata (match-only engine)"};
        const uint32_t Cn = m.symbols.n_classes, PAIR = m.symbols.pair_hi_class;

        // ---- joint alphabet: (DFA class, capture class) pairs that occur, then the PAIR_HI variants
        std::map<std::pair<uint32_t, uint32_t>, uint16_t> jid;
        std::vector<std::pair<uint32_t, uint32_t>> jdef;
        auto joint = [&](uint32_t d, uint32_t c) -> uint16_t {
            auto it = jid.find({d, c});
            if (it != jid.end()) return it->second;
            if (jdef.size() >= 0xFFF0) throw Refuse{"too many joint symbol classes"};
            jid.emplace(std::make_pair(d, c), static_cast<uint16_t>(jdef.size()));
            jdef.push_back({d, c});
            return static_cast<uint16_t>(jdef.size() - 1);
        };
        A.jcls.resize(65536);
        for (uint32_t u = 0; u < 65536; ++u) A.jcls[u] = joint(D.classmap[u], m.symbols.classmap[u]);
        {
            const size_t base = jdef.size();
            std::vector<uint16_t> pair_of(base);
            for (size_t j = 0; j < base; ++j) pair_of[j] = static_cast<uint16_t>(j);
            for (uint32_t u = 0xD800; u < 0xDC00; ++u) pair_of[A.jcls[u]] = joint(D.classmap[u], PAIR);
            A.pair_of = pair_of;
            A.pair_of.resize(jdef.size());
            for (size_t j = base; j < jdef.size(); ++j) A.pair_of[j] = static_cast<uint16_t>(j);
        }
        const size_t J = jdef.size();
        A.n_jcls = static_cast<uint32_t>(J);

        // ---- which extractions can still be the first accepting index from each DFA state
        const size_t W = (E + 63) / 64;
        std::vector<uint64_t> reach(S * W, 0);
        for (size_t q = 0; q < S; ++q)
            if (D.accept_first[q] >= 0) reach[q * W + D.accept_first[q] / 64] |= 1ull << (D.accept_first[q] % 64);
        for (bool changed = true; changed;) {
            changed = false;
            for (size_t q = S; q-- > 0;)
                for (size_t c = 0; c < C; ++c) {
                    const int32_t t = D.trans[q * C + c];
                    if (t < 0) continue;
                    for (size_t w = 0; w < W; ++w) {
                        const uint64_t v = reach[q * W + w] | reach[static_cast<size_t>(t) * W + w];
                        if (v != reach[q * W + w]) {
                            reach[q * W + w] = v;
                            changed = true;
                        }
                    }
                }
        }
        auto alive = [&](size_t q, size_t e) { return (reach[q * W + e / 64] >> (e % 64)) & 1u; };

        // ---- the capture automata must be copy-free (see fused.hpp)
        for (size_t e = 0; e < E; ++e)
            for (uint16_t op : m.tdfas[e].ops)
                if ((op & 0xFF) != 0xFF) throw Refuse{strfmt("extraction #%zu: capture automaton copies tag registers", e)};

        // ---- product BFS. key = q, then (e, t_e) for the extractions alive at q, ascending e
        std::map<std::vector<uint32_t>, uint32_t> index;
        std::vector<std::vector<uint32_t>> states;
        std::map<std::vector<uint32_t>, uint32_t> slot_of;  // product op list [(e, oplist)...] -> op slot (1..)
        std::vector<std::vector<uint32_t>> slot_def{{}};
        std::vector<uint32_t> trans;                         // raw product, [P*J]
        {
            std::vector<uint32_t> k0{0};
            for (size_t e = 0; e < E; ++e)
                if (alive(0, e)) {
                    k0.push_back(static_cast<uint32_t>(e));
                    k0.push_back(0);
                }
            index.emplace(k0, 0);
            states.push_back(std::move(k0));
        }
        for (size_t p = 0; p < states.size(); ++p) {
            const std::vector<uint32_t> cur = states[p];
            const size_t q = cur[0];
            for (size_t j = 0; j < J; ++j) {
                const int32_t q2 = D.trans[q * C + jdef[j].first];
                if (q2 < 0) {
                    trans.push_back(0xFFFFu);
                    continue;
                }
                const uint32_t c = jdef[j].second;
                std::vector<uint32_t> key{static_cast<uint32_t>(q2)}, ops;
                for (size_t i = 1; i < cur.size(); i += 2) {
                    const uint32_t e = cur[i], t = cur[i + 1];
                    if (!alive(static_cast<size_t>(q2), e)) continue;
                    uint32_t t2 = kDeadT;
                    if (t != kDeadT) {
                        const Tdfa& T = m.tdfas[e];
                        const uint32_t ent = T.trans[static_cast<size_t>(t) * Cn + c];
                        if ((ent & 0xFFFFu) != 0xFFFFu) {
                            t2 = ent & 0xFFFFu;
                            const uint32_t ol = ent >> 16;
                            if (T.op_off[ol + 1] > T.op_off[ol]) {
                                ops.push_back(e);
                                ops.push_back(ol);
                            }
                        }
                    }
                    key.push_back(e);
                    key.push_back(t2);
                }
                uint32_t slot = 0;
                if (!ops.empty()) {
                    auto it = slot_of.find(ops);
                    if (it == slot_of.end()) {
                        if (slot_def.size() > max_op_slots) throw Refuse{"too many distinct register command lists"};
                        it = slot_of.emplace(ops, static_cast<uint32_t>(slot_def.size())).first;
                        slot_def.push_back(ops);
                    }
                    slot = it->second;
                }
                auto it = index.find(key);
                if (it == index.end()) {
                    if (states.size() >= max_states) throw Refuse{"product automaton exceeds the state limit"};
                    it = index.emplace(key, static_cast<uint32_t>(states.size())).first;
                    states.push_back(std::move(key));
                }
                trans.push_back(it->second | (slot << 16));
            }
        }
        const size_t P = states.size();

        // ---- writers of each capture register: (e, r) -> op slots
        std::map<std::pair<uint32_t, uint32_t>, std::vector<uint32_t>> writers;
        for (uint32_t s = 1; s < slot_def.size(); ++s)
            for (size_t i = 0; i < slot_def[s].size(); i += 2) {
                const uint32_t e = slot_def[s][i], ol = slot_def[s][i + 1];
                const Tdfa& T = m.tdfas[e];
                for (uint32_t o = T.op_off[ol]; o < T.op_off[ol + 1]; ++o) {
                    auto& w = writers[{e, static_cast<uint32_t>(T.ops[o] >> 8)}];
                    if (std::find(w.begin(), w.end(), s) == w.end()) w.push_back(s);
                }
            }

        // ---- outcomes
        std::map<std::vector<uint32_t>, uint32_t> outcome_id;
        A.outcomes.push_back({-1, 0});
        outcome_id.emplace(std::vector<uint32_t>{0xFFFFFFFFu}, 0);
        std::vector<uint32_t> outcome_of(P);
        for (size_t p = 0; p < P; ++p) {
            const std::vector<uint32_t>& st = states[p];
            const int32_t e = D.accept_first[st[0]];
            if (e < 0) {
                outcome_of[p] = 0;
                continue;
            }
            uint32_t t = kDeadT;
            for (size_t i = 1; i < st.size(); i += 2)
                if (st[i] == static_cast<uint32_t>(e)) t = st[i + 1];
            const Tdfa& T = m.tdfas[static_cast<size_t>(e)];
            std::vector<uint32_t> key;
            if (t == kDeadT || !T.accepting[t]) {
                key = {0xFFFFFFFEu, static_cast<uint32_t>(e)};
            } else {
                key = {static_cast<uint32_t>(e)};
                for (uint32_t k = 0; k < T.n_slots; ++k) key.push_back(T.fin[static_cast<size_t>(t) * T.n_slots + k]);
            }
            auto it = outcome_id.find(key);
            if (it == outcome_id.end()) {
                FusedAutomaton::Outcome o{};
                if (key[0] == 0xFFFFFFFEu) {
                    o.ext_code = -2 - e;
                } else {
                    o.ext_code = e;
                    o.res_off = static_cast<uint32_t>(A.res.size());
                    for (uint32_t k = 0; k < T.n_slots; ++k) {
                        const uint32_t f = key[1 + k];
                        uint32_t packed = 0;
                        if (f == 0xFE) {
                            packed = FusedAutomaton::kLenSlot;
                        } else if (f != 0xFF) {
                            auto w = writers.find({static_cast<uint32_t>(e), f});
                            if (w != writers.end()) {
                                if (w->second.size() > 4) throw Refuse{"a group boundary has more than 4 writers"};
                                for (size_t i = 0; i < w->second.size(); ++i) packed |= w->second[i] << (8 * i);
                            }
                        }
                        A.res.push_back(packed);
                    }
                }
                if (A.outcomes.size() >= 4096) throw Refuse{"too many distinct outcomes"};
                it = outcome_id.emplace(key, static_cast<uint32_t>(A.outcomes.size())).first;
                A.outcomes.push_back(o);
            }
            outcome_of[p] = it->second;
        }

        // ---- Moore minimisation: states are equivalent iff same outcome and, per symbol, same op slot and equivalent
        //      successor (block P == dead)
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
            std::vector<uint32_t> np(P + 1), key(J + 1);
            for (size_t p = 0; p <= P; ++p) {
                key[0] = part[p];
                for (size_t j = 0; j < J; ++j) {
                    const uint32_t ent = p == P ? 0xFFFFu : trans[p * J + j];
                    const uint32_t nx = ent & 0xFFFFu;
                    key[j + 1] = nx == 0xFFFFu ? (part[P] | 0x80000000u) : (part[nx] | ((ent >> 16) << 16));
                }
                np[p] = sig.emplace(key, static_cast<uint32_t>(sig.size())).first->second;
            }
            part.swap(np);
            if (sig.size() == nblocks) break;
            nblocks = sig.size();
        }
        // a live state equivalent to "dead" (never reaches a non-MISS outcome... it still must end as MISS): the dead
        // block's outcome is MISS and all its successors are dead, so merging is exact
        const uint32_t dead_block = part[P];
        std::vector<int64_t> num(nblocks, -1);
        std::vector<size_t> rep;
        if (part[0] == dead_block) {
            // nothing can ever match: single start state whose every transition is dead
            A.n_states = 1;
            A.trans.assign(J, 0xFFFFu);
            A.outcome_of.assign(1, 0);
        } else {
            num[part[0]] = 0;
            rep.push_back(0);
            for (size_t i = 0; i < rep.size(); ++i)
                for (size_t j = 0; j < J; ++j) {
                    const uint32_t ent = trans[rep[i] * J + j];
                    if ((ent & 0xFFFFu) == 0xFFFFu) continue;
                    const uint32_t b = part[ent & 0xFFFFu];
                    if (b == dead_block || num[b] >= 0) continue;
                    num[b] = static_cast<int64_t>(rep.size());
                    rep.push_back(ent & 0xFFFFu);
                }
            A.n_states = static_cast<uint32_t>(rep.size());
            A.trans.resize(rep.size() * J);
            A.outcome_of.resize(rep.size());
            for (size_t i = 0; i < rep.size(); ++i) {
                A.outcome_of[i] = outcome_of[rep[i]];
                for (size_t j = 0; j < J; ++j) {
                    const uint32_t ent = trans[rep[i] * J + j];
                    const uint32_t nx = ent & 0xFFFFu;
                    if (nx == 0xFFFFu || part[nx] == dead_block) A.trans[i * J + j] = 0xFFFFu;
                    else A.trans[i * J + j] = static_cast<uint32_t>(num[part[nx]]) | (ent & 0xFFFF0000u);
                }
            }
        }
        A.n_op_slots = static_cast<uint32_t>(slot_def.size() - 1);
        A.available = true;
    } catch (const Refuse& r) {
        A = FusedAutomaton{};
        A.why_not = r.why;
    }
    return A;
}

void finalize_device_model(DeviceModel& m, const FusedAutomaton& fused) {
    if (fused.available) return;
    for (Tdfa& t : m.tdfas) minimise_tdfa(t);
}

}  // namespace gorp
