#include "walktables.hpp"

#include <algorithm>
#include <cstring>

namespace gorp {

DfaWalkTable build_dfawalk_table(const DeviceModel& m) {
    DfaWalkTable t;
    const size_t S = m.dfa.n_states, C = m.dfa.n_classes, E = m.n_groups.size();
    // A table that fits shared memory gets an odd column count (the rows of lanes that read the same column then fall
    // into different banks); a larger one is read through L1/L2 and gets 32-byte aligned rows whose first sector holds
    // the 16 hottest columns, so a warp-wide lookup touches fewer sectors and the hot set is smaller.
    const size_t NL = C, dscan = S, fin_base = S + 16, R = fin_base + 1 + E;
    const bool big = R * ((C + 1) | 1) * 2 > 200 * 1024;
    const size_t K = big ? ((C + 1 + 15) / 16) * 16 : ((C + 1) | 1);
    if (R > 0xFFFF || K * 2 > 0xFFFF) return t;
    std::vector<uint32_t> weight(C, 0), col_of(C + 1), cls_of(C + 1);
    for (uint32_t u = 0; u < 128; ++u) {  // a static weight per ASCII character: how often log text hits a class
        if (u == 0x0A) continue;
        uint32_t w = 1;
        if (u >= 'a' && u <= 'z') w = 8;
        else if ((u >= '0' && u <= '9') || u == ' ') w = 6;
        else if (u >= 'A' && u <= 'Z') w = 3;
        else if (u && std::strchr(".-_/:=[](),'\"", static_cast<int>(u))) w = 2;
        weight[m.dfa.classmap[u]] += w;
    }
    for (size_t k = 0; k <= C; ++k) cls_of[k] = static_cast<uint32_t>(k);
    std::stable_sort(cls_of.begin(), cls_of.begin() + C, [&](uint32_t a, uint32_t b) { return weight[a] > weight[b]; });
    for (size_t k = 0; k <= C; ++k) col_of[cls_of[k]] = static_cast<uint32_t>(k);  // class -> column; '\n' (id C) stays last
    t.rows.assign(((R * K + 7) / 8) * 8, 0);
    for (size_t r = 0; r < R; ++r)
        for (size_t k = 0; k < K; ++k) {
            size_t nx;
            if (k > NL) {
                nx = r;  // padding column, never addressed
            } else if (r < S) {
                if (k == NL) {
                    nx = fin_base + 1 + m.dfa.accept_first[r];
                } else {
                    const int32_t to = m.dfa.trans[r * C + cls_of[k]];
                    nx = to < 0 ? dscan : static_cast<size_t>(to);
                }
            } else if (r == dscan) {
                nx = k == NL ? fin_base : dscan;
            } else if (r < fin_base) {
                nx = r == S + 1 ? 0 : r - 1;  // SKIP chain (swallows any unit, '\n' included)
            } else {
                nx = r;  // FIN: absorbing
            }
            t.rows[r * K + k] = static_cast<uint16_t>(nx);
        }
    t.cls128.resize(128);
    t.xcls.resize(65536);
    for (size_t u = 0; u < 65536; ++u) t.xcls[u] = static_cast<uint16_t>(col_of[m.dfa.classmap[u]]);
    for (size_t u = 0; u < 128; ++u) t.cls128[u] = static_cast<uint16_t>(2 * (u == 0x0A ? NL : col_of[m.dfa.classmap[u]]));
    t.n_rows = static_cast<uint32_t>(R);
    t.K = static_cast<uint32_t>(K);
    t.n_states = static_cast<uint32_t>(S);
    t.n_classes = static_cast<uint32_t>(C);
    t.fin_base = static_cast<uint32_t>(fin_base);
    t.available = true;
    return t;
}

DfaWalkTable build_dfawalk_table_cut(const DeviceModel& m, const std::vector<int32_t>& cut_of_state) {
    DfaWalkTable t;
    const size_t S0 = m.dfa.n_states, C = m.dfa.n_classes, E = m.n_groups.size();
    if (cut_of_state.size() != S0) return t;
    // head states: reachable from the start without passing a cut state, renumbered in BFS order
    std::vector<int32_t> id(S0, -1);
    std::vector<uint32_t> order{0};
    id[0] = 0;
    for (size_t i = 0; i < order.size(); ++i)
        for (size_t c = 0; c < C; ++c) {
            const int32_t to = m.dfa.trans[static_cast<size_t>(order[i]) * C + c];
            if (to < 0 || cut_of_state[to] >= 0 || id[to] >= 0) continue;
            id[to] = static_cast<int32_t>(order.size());
            order.push_back(static_cast<uint32_t>(to));
        }
    const size_t S = order.size();
    const size_t NL = C, fin_base = S + 16, R = fin_base + 1 + E;
    const size_t K = (C + 1) | 1;  // shared memory: odd column count (see build_dfawalk_table)
    if (R > 0xFFFF || K * 2 > 0xFFFF) return t;
    t.rows.assign(((R * K + 7) / 8) * 8, 0);
    for (size_t r = 0; r < R; ++r)
        for (size_t k = 0; k < K; ++k) {
            size_t nx;
            if (k > NL) {
                nx = r;  // padding column, never addressed
            } else if (r < S) {
                const size_t q = order[r];
                if (k == NL) {
                    nx = fin_base + 1 + m.dfa.accept_first[q];
                } else {
                    const int32_t to = m.dfa.trans[q * C + k];
                    if (to < 0) nx = fin_base;                                        // dead: MISS at once
                    else if (cut_of_state[to] >= 0) nx = fin_base + 1 + cut_of_state[to];  // candidate: the tail decides
                    else nx = static_cast<size_t>(id[to]);
                }
            } else if (r == S) {
                nx = fin_base;  // DEADSCAN is not used by this variant
            } else if (r < fin_base) {
                nx = r == S + 1 ? 0 : r - 1;  // SKIP chain
            } else {
                nx = r;  // FIN: absorbing
            }
            t.rows[r * K + k] = static_cast<uint16_t>(nx);
        }
    t.cls128.resize(128);
    t.xcls.resize(65536);
    for (size_t u = 0; u < 65536; ++u) t.xcls[u] = m.dfa.classmap[u];
    for (size_t u = 0; u < 128; ++u) t.cls128[u] = static_cast<uint16_t>(2 * (u == 0x0A ? NL : m.dfa.classmap[u]));
    t.n_rows = static_cast<uint32_t>(R);
    t.K = static_cast<uint32_t>(K);
    t.n_states = static_cast<uint32_t>(S);
    t.n_classes = static_cast<uint32_t>(C);
    t.fin_base = static_cast<uint32_t>(fin_base);
    t.available = true;
    return t;
}

TailImage build_tail_image(const TailSet& T, uint32_t span_stride) {
    TailImage I;
    if (!T.any) return I;
    const size_t E = T.tails.size();
    I.width = T.width + 2;  // an odd number of 32-bit words per row: the rows of lanes in different states spread over the banks
    I.row_bytes = I.width * 2;
    I.span_stride = span_stride;
    I.ext.assign(E, TailImageExt{});
    for (size_t e = 0; e < E; ++e) {
        const TailAutomaton& A = T.tails[e];
        if (!A.available) continue;
        const uint32_t S = A.n_states, fin_base = S + 15, n_out = static_cast<uint32_t>(A.outcomes.size()), rows = fin_base + n_out;
        if (rows > 1023 || A.n_op_slots + 1 > 63) continue;  // 10-bit row ids, 6-bit slots
        size_t n_multi = 0;  // group boundaries with several writers (kernels/tailwalk.cu keeps at most 64 of them per table)
        for (const FusedAutomaton::Outcome& oc : A.outcomes)
            if (oc.ext_code >= 0)
                for (uint32_t k = 0; k < A.n_boundaries && k < span_stride; ++k) n_multi += A.res[oc.res_off + k] >= 256u ? 1 : 0;
        if (n_multi > 64) continue;
        while (I.image.size() % 8) I.image.push_back(0);
        TailImageExt& x = I.ext[e];
        x.tab_off = static_cast<uint32_t>(I.image.size() * 2);
        x.n_states = S;
        x.fin_base = fin_base;
        x.n_outcomes = n_out;
        x.res_off = static_cast<uint32_t>(I.res.size());
        x.oext_off = static_cast<uint32_t>(I.oext.size());
        x.init_off = static_cast<uint32_t>(I.init_slots.size());
        x.n_init = static_cast<uint32_t>(A.init_slots.size());
        x.n_slots = A.n_op_slots + 1;
        x.available = 1;
        const size_t base = I.image.size();
        I.image.resize(base + static_cast<size_t>(rows) * I.width);
        auto put = [&](uint32_t r, uint32_t k, uint32_t next_row, uint32_t slot) {
            I.image[base + static_cast<size_t>(r) * I.width + k] = static_cast<uint16_t>((next_row << 6) | slot);
        };
        for (uint32_t r = 0; r < rows; ++r)
            for (uint32_t k = 0; k < I.width; ++k) {
                if (r < S) {
                    if (k == 0x0A) {
                        put(r, k, fin_base + A.outcome_of[r], 0);
                    } else {
                        const uint32_t ent = k < T.width ? A.trans[static_cast<size_t>(r) * T.width + k] : 0xFFFFu;
                        if ((ent & 0xFFFFu) == 0xFFFFu) put(r, k, fin_base, 0);  // dead (or a padding column): MISS
                        else put(r, k, ent & 0xFFFFu, ent >> 16);
                    }
                } else if (r < fin_base) {
                    put(r, k, r == S ? 0u : r - 1, 0);  // SKIP chain
                } else {
                    put(r, k, r, 0);  // outcome rows: absorbing
                }
            }
        for (uint32_t o = 0; o < n_out; ++o) {
            const FusedAutomaton::Outcome& oc = A.outcomes[o];
            I.oext.push_back(oc.ext_code);
            for (uint32_t k = 0; k < span_stride; ++k) {
                uint32_t rec = 0;
                if (oc.ext_code >= 0 && k < A.n_boundaries) rec = A.res[oc.res_off + k];  // 2 * groups recipes per MATCH outcome
                I.res.push_back(rec);
            }
        }
        for (uint32_t s : A.init_slots) I.init_slots.push_back(static_cast<uint8_t>(s));
        I.max_table_bytes = std::max<uint32_t>(I.max_table_bytes, (rows * I.row_bytes + 15u) & ~15u);
        I.max_slots = std::max(I.max_slots, x.n_slots);
        I.available = true;
    }
    while (I.image.size() % 8) I.image.push_back(0);
    I.image.resize(I.image.size() + 16, 0);
    I.init_slots.push_back(0);
    return I;
}

CapImage build_cap_image(const DeviceModel& m, size_t max_extractions) {
    CapImage img;
    const size_t E = m.n_groups.size();
    if (E > max_extractions || m.tdfas.size() != E) return img;
    // K = classes + '\n' column, padded to an odd count (shared-memory banks, as for the DFA table)
    const uint32_t Cn = m.symbols.n_classes, NL = Cn, K = (Cn + 1) | 1, row_bytes = K * 4;
    uint32_t max_regs = 0;
    for (size_t e = 0; e < E; ++e) max_regs = std::max(max_regs, m.tdfas[e].n_regs);
    if (max_regs >= 63) return img;  // the register slot is 6 bits of the entry
    img.cls128.resize(128);
    for (uint32_t u = 0; u < 128; ++u) img.cls128[u] = (u == 0x0A ? NL : m.symbols.classmap[u]) * 4;
    img.ext.resize(E);
    std::vector<uint32_t>& image = img.image;
    for (size_t e = 0; e < E; ++e) {
        const Tdfa& t = m.tdfas[e];
        const uint32_t Sx = t.n_states, rows = 2 * Sx + 17;
        if (static_cast<uint64_t>(rows) * row_bytes >= (1u << 26) || (image.size() + static_cast<size_t>(rows) * K) * 4 >= (1ull << 31)) return img;
        while (image.size() % 4) image.push_back(0);  // 16-byte aligned tables (copied to shared memory with 128-bit loads)
        img.ext[e] = {static_cast<uint32_t>(image.size() * 4), row_bytes, Sx, (Sx + 15) * row_bytes, (Sx + 16) * row_bytes, (Sx + 17) * row_bytes};
        const size_t base = image.size();
        image.resize(base + static_cast<size_t>(rows) * K);
        auto put = [&](uint32_t r, uint32_t k, uint32_t next_row, uint32_t slot) {
            image[base + static_cast<size_t>(r) * K + k] = ((next_row * row_bytes) << 6) | slot;
        };
        for (uint32_t r = 0; r < rows; ++r)
            for (uint32_t k = 0; k < K; ++k) {
                if (k > NL) { put(r, k, r, max_regs); continue; }  // padding column, never addressed
                if (r < Sx) {
                    if (k == NL) { put(r, k, Sx + 17 + r, max_regs); continue; }
                    const uint32_t ent = t.trans[static_cast<size_t>(r) * Cn + k];
                    const uint32_t nx = ent & 0xFFFFu, ol = ent >> 16;
                    if (nx == 0xFFFFu) { put(r, k, Sx + 15, max_regs); continue; }
                    const uint32_t o0 = t.op_off[ol], o1 = t.op_off[ol + 1];
                    if (o1 == o0) put(r, k, nx, max_regs);
                    else if (o1 - o0 == 1 && (t.ops[o0] & 0xFF) == 0xFF) put(r, k, nx, t.ops[o0] >> 8);
                    else put(r, k, Sx + 16, max_regs);  // SLOW: replayed through the general tables
                } else if (r < Sx + 15) {
                    put(r, k, r == Sx ? 0u : r - 1, max_regs);  // SKIP chain
                } else {
                    put(r, k, r, max_regs);  // DEAD / SLOW / FRZ: absorbing
                }
            }
    }
    image.resize(image.size() + 8, 0);  // the last table may be read 16 bytes at a time
    img.K = K;
    img.n_regs = max_regs;
    img.available = true;
    return img;
}

}  // namespace gorp
