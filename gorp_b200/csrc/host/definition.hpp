// Definition front-end: `.grp` text -> per extraction (name, extractor names, automaton-dialect regex,
// JDK-dialect regex, raw `append` JSON). Load-time only; mirrors the observable behaviour of the
// reference's DefinitionReader.read() + Gorp.construct() string generation:
//   gorp-core/src/main/java/com/salesforce/gorp/DefinitionReader.java:74-640
//   .../model/CookedDefinitions.java:57-475, .../util/RegexHelper.java:20-237, .../Gorp.java:50-129
#pragma once
#include "common.hpp"

namespace gorp {

struct ExtractionStrings {
    ustring name;
    std::vector<ustring> extractor_names;  // group-number order
    ustring automaton_regex;               // brics dialect (Gorp.java:68)
    ustring jdk_regex;                     // java.util.regex dialect (Gorp.java:70)
    std::string append_json;               // merged raw JSON object text of all `append` lines ("" if none)
};

// RegexHelper.java:20-70 / :79-201 / :210-237
ustring quote_literal_as_regexp(const ustring& text);
ustring massage_regexp_for_automaton(const ustring& pattern);  // throws std::invalid_argument
ustring massage_regexp_for_jdk(const ustring& pattern);

// DefinitionReader.reader(String).read() up to the regex strings. Throws DefinitionParseError.
std::vector<ExtractionStrings> read_definition(const ustring& text, const std::string& source_ref = "<input string>");

}  // namespace gorp
