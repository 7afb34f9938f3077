// Shared host-side types for libgorpcuda (pure C++17, no CUDA).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace gorp {

// Error categories surfaced through the C ABI (include/gorp_cuda.h).
struct DefinitionParseError : std::runtime_error {  // ~ DefinitionParseException (reference)
    using std::runtime_error::runtime_error;
};
struct UnsupportedError : std::runtime_error {  // legal for the reference, refused by the GPU path
    using std::runtime_error::runtime_error;
};
struct BlobError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

using ustring = std::u16string;  // Java String == UTF-16 code units

ustring utf8_to_utf16(const char* s, size_t n);
std::string utf16_to_utf8(const ustring& s);

std::string strfmt(const char* fmt, ...);

}  // namespace gorp
