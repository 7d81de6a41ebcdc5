// oracle/ref_coders_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" driver over the REFERENCE's own legacy symbol packers, compiled from where they lie
// (/root/reference/coders/SymbolsPackingFacility.cpp, coders/VarLenDNACoder.cpp): row a14 of SURVEY.md 8.  No reference
// run reaches these classes, so this driver is what pins the oracle restatement (oracle/mbgc_oracle.c:
// mbo_symbols_*, mbo_varlen_*) and the CUDA kernels (mbgc_b200/csrc/coders.cu).  No algorithm of its own.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "coders/SymbolsPackingFacility.h"
#include "coders/VarLenDNACoder.h"

extern "C" {

// SymbolsPackingFacility(symbols).packSequence(src, n, dst); returns the packed length
uint64_t refc_symbols_pack(const char* symbols, int nSymbols, const char* src, uint64_t n, uint8_t* dst) {
    std::vector<char> list(symbols, symbols + nSymbols);
    PgIndex::SymbolsPackingFacility f(list);
    return f.packSequence(src, n, dst);
}

// reverseSequence(sequence, pos, length, destPtr)
void refc_symbols_unpack(const char* symbols, int nSymbols, const uint8_t* packed, uint64_t pos, uint64_t length, char* dst) {
    std::vector<char> list(symbols, symbols + nSymbols);
    PgIndex::SymbolsPackingFacility f(list);
    if (length) f.reverseSequence(packed, pos, length, dst);
}

// the static code book `id` (0..2) as writeBook emits it; returns its length (without the final 0)
uint64_t refc_varlen_book(int id, char* out /* >= 2048 bytes */) {
    PgHelpers::VarLenDNACoder c(static_cast<uint8_t>(id));
    size_t n = 0;
    c.writeBook(reinterpret_cast<unsigned char*>(out), n);
    return n - 1;
}

uint64_t refc_varlen_encode(const char* codeBook, const uint8_t* src, uint64_t n, uint8_t* dst) {
    PgHelpers::VarLenDNACoder c{std::string(codeBook)};
    size_t outLen = 0;
    c.encode(dst, outLen, src, n);
    return outLen;
}

void refc_varlen_decode(const char* codeBook, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t expLen) {
    PgHelpers::VarLenDNACoder c{std::string(codeBook)};
    c.decode(dst, expLen, const_cast<unsigned char*>(src), n);
}

}  // extern "C"
