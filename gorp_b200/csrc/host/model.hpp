// Compiled definition: the host-side product of the definition compiler and the content of the DfaExport blob.
//
// Blob layout (little-endian), the serialised form a Java-side DfaExport produces from
// Automata._alphabet/_transitions/_accept (reference autom/Automata.java:23-26) and
// CookedExtraction.getName()/getRegexpSource() (reference model/CookedExtraction.java:38,46):
//
//   header    : "GORPDFA1", u32 version(=1), u32 n_states S, u32 n_classes C, u32 n_extractions E,
//               u32 flags(=0), u32 reserved, u64 fnv1a64(body)
//   classmap  : u16[65536]
//   trans     : i32[S*C]            row-major, -1 = dead, start state 0
//   acc_first : i32[S]
//   acc_off   : u32[S+1], acc_list : i32[acc_off[S]]
//   E x       : u32 n_groups, str name, str automaton_regex, str jdk_regex, u32 n_names, str names[n_names],
//               u32 json_len, u8 json[json_len] (+pad to 4)        str = u32 n_units, u16[n_units] (+pad to 4)
#pragma once
#include "automata.hpp"
#include "capture.hpp"
#include "definition.hpp"

namespace gorp {

struct CompiledExtraction {
    uint32_t n_groups = 0;
    ExtractionStrings strings;
};

struct CompiledDefinition {
    DfaTables dfa;
    std::vector<CompiledExtraction> extractions;
};

// DefinitionReader.read() + Gorp.construct(): definition text -> tables + regex strings (no JVM needed).
CompiledDefinition compile_definition(const ustring& grp_text);
// PolyMatcher.create(patterns) alone (reference autom/PolyMatcher.java:68-84): raw brics-dialect regexes.
DfaTables compile_patterns(const std::vector<ustring>& automaton_regexes);

std::vector<uint8_t> serialize_blob(const CompiledDefinition& d);
CompiledDefinition parse_blob(const void* data, size_t len);  // throws BlobError

// Everything the device needs, derived from a CompiledDefinition at engine creation.
struct DeviceModel {
    CompactDfa dfa;
    SymbolClasses symbols;
    std::vector<CaptureProgram> programs;
    std::vector<Tdfa> tdfas;
    std::vector<uint32_t> n_groups;
    // extractions whose capture automaton could not be determinised within the limits: `tdfas[e]` is a placeholder that
    // accepts nothing and `pike[e]` holds the tables of the simulating Pike-VM pass that decides their lines
    std::vector<uint8_t> pike_only;
    std::vector<PikeTables> pike;
};
DeviceModel build_device_model(const CompiledDefinition& d);

}  // namespace gorp
