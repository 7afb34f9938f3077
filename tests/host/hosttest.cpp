// TEST-ONLY library (never part of libgorpcuda.so, never shipped): interprets the tables that the product's
// definition compiler produces (CompactDfa + per-extraction TDFA) on the CPU, so that the host-side compile
// logic can be checked against the oracle without a GPU. The loops mirror kernels/kernels.cu (K2, K4).
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "../../gorp_b200/csrc/host/model.hpp"

using namespace gorp;

extern "C" int ht_run(const void* blob, size_t len, const uint16_t* text, const int64_t* starts, const int64_t* ends, int64_t n,
                      int32_t* ext, int32_t* spans, int stride, uint32_t* stats, char* err, int errlen) {
    try {
        CompiledDefinition def = parse_blob(blob, len);
        DeviceModel m = build_device_model(def);
        if (stats) {
            stats[0] = m.dfa.n_states;
            stats[1] = m.dfa.n_classes;
            stats[2] = m.symbols.n_classes;
            uint32_t ts = 0, tr = 0;
            for (auto& t : m.tdfas) {
                ts = std::max(ts, t.n_states);
                tr = std::max(tr, t.n_regs);
            }
            stats[3] = ts;
            stats[4] = tr;
        }
        const uint32_t C = m.dfa.n_classes;
        for (int64_t l = 0; l < n; ++l) {
            const uint16_t* u = text + starts[l];
            const int64_t L = ends[l] - starts[l];
            int32_t st = 0;
            for (int64_t i = 0; i < L && st >= 0; ++i) st = m.dfa.trans[static_cast<size_t>(st) * C + m.dfa.classmap[u[i]]];
            int32_t e = st < 0 ? -1 : m.dfa.accept_first[st];
            ext[l] = e;
            for (int k = 0; k < stride; ++k) spans[l * stride + k] = -1;
            if (e < 0) continue;
            const Tdfa& t = m.tdfas[e];
            std::vector<int32_t> regs(t.n_regs + 1, -7);
            uint32_t s = 0;
            bool ok = true;
            for (int64_t i = 0; i < L; ++i) {
                uint32_t k = m.symbols.classmap[u[i]];
                if ((u[i] & 0xFC00) == 0xD800 && i + 1 < L && (u[i + 1] & 0xFC00) == 0xDC00) k = m.symbols.pair_hi_class;
                uint32_t ent = t.trans[static_cast<size_t>(s) * t.n_classes + k];
                uint32_t nx = ent & 0xFFFF;
                if (nx == 0xFFFF) {
                    ok = false;
                    break;
                }
                uint32_t ol = ent >> 16;
                for (uint32_t q = t.op_off[ol]; q < t.op_off[ol + 1]; ++q) {
                    uint32_t op = t.ops[q], src = op & 0xFF;
                    regs[op >> 8] = src == 0xFF ? static_cast<int32_t>(i) : regs[src];
                }
                s = nx;
            }
            if (ok) ok = t.accepting[s] != 0;
            if (!ok) {
                ext[l] = -2 - e;
                continue;
            }
            for (uint32_t k = 0; k < t.n_slots && static_cast<int>(k) < stride; ++k) {
                uint8_t f = t.fin[static_cast<size_t>(s) * t.n_slots + k];
                spans[l * stride + k] = f == 0xFF ? -1 : (f == 0xFE ? static_cast<int32_t>(L) : regs[f]);
            }
        }
        return 0;
    } catch (const std::exception& e) {
        std::snprintf(err, errlen, "%s", e.what());
        return -1;
    }
}
