// mbgc_b200/csrc/coders.h -- host-visible launchers of coders.cu (row a14 of SURVEY.md 8: SymbolsPackingFacility,
// VarLenDNACoder of /root/reference/coders/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mbgc_b200 {

// order[256]: index of a byte in the alphabet, -1 = not a symbol; perElement = symbols per packed byte
int launch_sym_pack(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t nOut, const int8_t* order, int symbolsCount,
                    int perElement, uint32_t* bad, cudaStream_t st);
// reverse[value * perElement + j] = j-th symbol of a packed value
int launch_sym_unpack(const uint8_t* packed, uint64_t pos, uint64_t length, uint8_t* dst, const uint8_t* reverse, int perElement,
                      cudaStream_t st);

// book: 256 words of 5 bytes (up to 4 symbols, zero padded); lut: 2^24 bytes
int launch_vl_build_lut(const uint8_t* book, int nCodes, uint8_t* lut, cudaStream_t st);
uint64_t vl_chunks(uint64_t n);
uint64_t vl_blocks(uint64_t n);
uint64_t vl_dec_blocks(uint64_t n);
// bad[0] bit 1: unexpected symbol on the parse path
int launch_vl_encode(const uint8_t* src, uint64_t n, const uint8_t* lut, uint4* chunkMap, uint4* blockMap, uint64_t* blockEntry,
                     uint64_t* total, uint8_t* dst, uint64_t dstCap, uint32_t* bad, cudaStream_t st);
int launch_vl_decode(const uint8_t* codes, uint64_t n, const uint8_t* book, uint32_t* blockLen, uint64_t* blockOff, uint64_t* total,
                     uint8_t* dst, uint64_t dstCap, cudaStream_t st);

}  // namespace mbgc_b200
