// mbgc_b200/csrc/coders.cpp -- host side of row a14: SymbolsPackingFacility::packSequence / reverseSequence and
// VarLenDNACoder::encode / decode (/root/reference/coders/) on the device.  See coders.cu.
#include "coders.h"

#include <cstring>
#include <string>
#include <vector>

#include "../../include/mbgc_b200.h"
#include "matcher.hpp"

namespace mbgc_b200 {

namespace {

struct DevBuf {  // scratch device memory of one call
    void* p = nullptr;
    explicit DevBuf(uint64_t n) {
        if (cudaMalloc(&p, n ? n : 1) != cudaSuccess) {
            cudaGetLastError();
            throw MatcherError(MBGCG_E_NOMEM, "cudaMalloc(coder scratch) failed");
        }
    }
    ~DevBuf() { cudaFree(p); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

void check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw MatcherError(MBGCG_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// SymbolsPackingFacility::maxSymbolsPerElement (coders/SymbolsPackingFacility.cpp:102-111): the largest k with N^k <= 256
int symbols_per_element(int n) {
    int k = 0;
    uint64_t pw = 1;
    while (pw * n <= 256) {
        pw *= n;
        k++;
    }
    return k;
}

// the code book as initUsing reads it (coders/VarLenDNACoder.cpp:10-27): words separated by '\n', at most 256, at most 4
// symbols each (strncpy into 5-byte slots), the first a single symbol
int parse_book(const char* codes, uint8_t book[256 * 5]) {
    std::memset(book, 0, 256 * 5);
    const std::string s(codes);
    size_t pos = 0;
    int counter = 0;
    for (;;) {
        const size_t nl = s.find('\n', pos);
        const size_t end = nl == std::string::npos ? s.size() : nl;
        if (counter > 255) throw MatcherError(MBGCG_E_PARAM, "Too many codes in the var-len coder book!");
        const size_t len = end - pos;
        if (len > 4) throw MatcherError(MBGCG_E_PARAM, "a var-len code word has at most 4 symbols");
        std::memcpy(book + counter * 5, s.data() + pos, len);
        counter++;
        if (nl == std::string::npos) break;
        pos = nl + 1;
    }
    if (book[0] == 0 || book[1] != 0) throw MatcherError(MBGCG_E_PARAM, "First code in the var-len coder book should be a single symbol!");
    return counter;
}

}  // namespace

uint64_t symbols_pack(B200SparseEMMatcher* m, const char* symbols, int nSymbols, const char* src, uint64_t n, uint8_t* dst) {
    if (nSymbols < 2 || nSymbols > 255) throw MatcherError(MBGCG_E_PARAM, "symbols_pack: between 2 and 255 symbols");
    check(cudaSetDevice(m->deviceIndex()), "cudaSetDevice");
    cudaStream_t st = static_cast<cudaStream_t>(m->streamHandle());
    const int per = symbols_per_element(nSymbols);
    const uint64_t nOut = (n + per - 1) / per;
    int8_t order[256];
    std::memset(order, -1, sizeof(order));
    for (int i = 0; i < nSymbols; i++) order[static_cast<uint8_t>(symbols[i])] = static_cast<int8_t>(i);
    DevBuf dSrc(n), dDst(nOut), dOrder(256), dBad(4);
    check(cudaMemcpyAsync(dSrc.p, src, n, cudaMemcpyHostToDevice, st), "H2D(symbols)");
    check(cudaMemcpyAsync(dOrder.p, order, 256, cudaMemcpyHostToDevice, st), "H2D(order)");
    check(cudaMemsetAsync(dBad.p, 0, 4, st), "memset");
    m->totalLaunches += launch_sym_pack(dSrc.as<uint8_t>(), n, dDst.as<uint8_t>(), nOut, dOrder.as<int8_t>(), nSymbols, per, dBad.as<uint32_t>(), st);
    uint32_t bad = 0;
    check(cudaMemcpyAsync(dst, dDst.p, nOut, cudaMemcpyDeviceToHost, st), "D2H(packed)");
    check(cudaMemcpyAsync(&bad, dBad.p, 4, cudaMemcpyDeviceToHost, st), "D2H(flag)");
    check(cudaStreamSynchronize(st), "symbols_pack sync");
    if (bad) throw MatcherError(MBGCG_E_PARAM, "Unexpected symbol: not in the packer's alphabet");
    return nOut;
}

void symbols_unpack(B200SparseEMMatcher* m, const char* symbols, int nSymbols, const uint8_t* packed, uint64_t pos, uint64_t length,
                    char* dst) {
    if (nSymbols < 2 || nSymbols > 255) throw MatcherError(MBGCG_E_PARAM, "symbols_unpack: between 2 and 255 symbols");
    check(cudaSetDevice(m->deviceIndex()), "cudaSetDevice");
    cudaStream_t st = static_cast<cudaStream_t>(m->streamHandle());
    const int per = symbols_per_element(nSymbols);
    uint64_t combos = 1;
    for (int j = 0; j < per; j++) combos *= nSymbols;
    std::vector<uint8_t> reverse(256 * per, static_cast<uint8_t>(symbols[0]));
    for (uint64_t v = 0; v < combos; v++) {  // value v = symbols in base N, first symbol most significant
        uint64_t x = v;
        for (int j = per - 1; j >= 0; j--) {
            reverse[v * per + j] = static_cast<uint8_t>(symbols[x % nSymbols]);
            x /= nSymbols;
        }
    }
    const uint64_t nPacked = (pos + length + per - 1) / per;
    DevBuf dPacked(nPacked), dDst(length), dRev(reverse.size());
    check(cudaMemcpyAsync(dPacked.p, packed, nPacked, cudaMemcpyHostToDevice, st), "H2D(packed)");
    check(cudaMemcpyAsync(dRev.p, reverse.data(), reverse.size(), cudaMemcpyHostToDevice, st), "H2D(reverse table)");
    m->totalLaunches += launch_sym_unpack(dPacked.as<uint8_t>(), pos, length, dDst.as<uint8_t>(), dRev.as<uint8_t>(), per, st);
    check(cudaMemcpyAsync(dst, dDst.p, length, cudaMemcpyDeviceToHost, st), "D2H(symbols)");
    check(cudaStreamSynchronize(st), "symbols_unpack sync");
}

uint64_t varlen_encode(B200SparseEMMatcher* m, const char* codeBook, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dstCap) {
    check(cudaSetDevice(m->deviceIndex()), "cudaSetDevice");
    cudaStream_t st = static_cast<cudaStream_t>(m->streamHandle());
    uint8_t book[256 * 5];
    const int nCodes = parse_book(codeBook, book);
    const uint64_t nChunks = vl_chunks(n), nBlocks = vl_blocks(n);
    DevBuf dSrc(n + 8), dLut(1u << 24), dBook(sizeof(book)), dChunk((nChunks + 1) * 16), dBlock((nBlocks + 1) * 16),
        dEntry((nBlocks + 1) * 8), dTotal(8), dDst(dstCap), dBad(8);
    check(cudaMemcpyAsync(dSrc.p, src, n, cudaMemcpyHostToDevice, st), "H2D(symbols)");
    check(cudaMemcpyAsync(dBook.p, book, sizeof(book), cudaMemcpyHostToDevice, st), "H2D(code book)");
    check(cudaMemsetAsync(dBad.p, 0, 8, st), "memset");
    int launches = launch_vl_build_lut(dBook.as<uint8_t>(), nCodes, dLut.as<uint8_t>(), st);
    launches += launch_vl_encode(dSrc.as<uint8_t>(), n, dLut.as<uint8_t>(), dChunk.as<uint4>(), dBlock.as<uint4>(), dEntry.as<uint64_t>(),
                                 dTotal.as<uint64_t>(), dDst.as<uint8_t>(), dstCap, dBad.as<uint32_t>(), st);
    m->totalLaunches += launches;
    uint64_t total = 0;
    uint32_t bad[2] = {0, 0};
    check(cudaMemcpyAsync(&total, dTotal.p, 8, cudaMemcpyDeviceToHost, st), "D2H(total)");
    check(cudaMemcpyAsync(bad, dBad.p, 8, cudaMemcpyDeviceToHost, st), "D2H(flag)");
    check(cudaStreamSynchronize(st), "varlen_encode sync");
    if (bad[0] & 2) throw MatcherError(MBGCG_E_PARAM, "ERROR: var-len encoder unexpected encoutered symbol!");
    if (total > dstCap) throw MatcherError(MBGCG_E_PARAM, "varlen_encode: destination too small");
    check(cudaMemcpyAsync(dst, dDst.p, total, cudaMemcpyDeviceToHost, st), "D2H(codes)");
    check(cudaStreamSynchronize(st), "varlen_encode sync");
    return total;
}

void varlen_decode(B200SparseEMMatcher* m, const char* codeBook, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t expLen) {
    check(cudaSetDevice(m->deviceIndex()), "cudaSetDevice");
    cudaStream_t st = static_cast<cudaStream_t>(m->streamHandle());
    uint8_t book[256 * 5];
    parse_book(codeBook, book);
    const uint64_t nb = vl_dec_blocks(n);
    DevBuf dSrc(n), dBook(sizeof(book)), dLen((nb + 1) * 4), dOff((nb + 1) * 8), dTotal(8), dDst(expLen);
    check(cudaMemcpyAsync(dSrc.p, src, n, cudaMemcpyHostToDevice, st), "H2D(codes)");
    check(cudaMemcpyAsync(dBook.p, book, sizeof(book), cudaMemcpyHostToDevice, st), "H2D(code book)");
    m->totalLaunches += launch_vl_decode(dSrc.as<uint8_t>(), n, dBook.as<uint8_t>(), dLen.as<uint32_t>(), dOff.as<uint64_t>(),
                                         dTotal.as<uint64_t>(), dDst.as<uint8_t>(), expLen, st);
    uint64_t total = 0;
    check(cudaMemcpyAsync(&total, dTotal.p, 8, cudaMemcpyDeviceToHost, st), "D2H(total)");
    check(cudaStreamSynchronize(st), "varlen_decode sync");
    if (total != expLen)  // VarLenDNACoder::decode, :121-125
        throw MatcherError(MBGCG_E_PARAM, "Unexpected decoded length: " + std::to_string(total) + " (expected " + std::to_string(expLen) + ").");
    check(cudaMemcpyAsync(dst, dDst.p, expLen, cudaMemcpyDeviceToHost, st), "D2H(symbols)");
    check(cudaStreamSynchronize(st), "varlen_decode sync");
}

}  // namespace mbgc_b200
