// mbgc_b200/csrc/ingest.cpp -- host side of the FASTA ingest: B200SparseEMMatcher::ingestFasta (matcher.hpp, ingest.cu).
//
// Mirrors what MBGC gets out of its reader for one file (/root/reference/utils/kseq.h:233-274 called in a loop,
// matching/MultipleGenomeMatchingProcessor.cpp:126-139 / 400-406): the records in file order (header line, sequence
// with the line ends removed), the file's DNA line length (KSEQ_DNA_LINE_LENGTH, :12-14) and the status that ended the
// loop (-1 end of file, -3 not FASTA, -4 DNA not well-formed).  The sequences stay on the device, laid out as a batch
// of texts for mbgcg_match_batch_device.
#include <algorithm>
#include <cstring>

#include "../../include/mbgc_b200.h"
#include "matcher.hpp"

namespace mbgc_b200 {

namespace {
constexpr uint64_t PAD = 64, TEXT_GAP = 64, TEXT_TAIL = 8192;  // = matcher.cpp
inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }
}  // namespace

void B200SparseEMMatcher::releaseIngest() {
    cudaFree(dFaIn);
    cudaFree(dFaWork);
    cudaFree(dFaText);
    if (hFaWork) cudaFreeHost(hFaWork);
    for (auto& e : faEv)
        if (e) cudaEventDestroy(e);
}

void B200SparseEMMatcher::ingestFasta(const char* bytes, uint64_t n) {
    check(cudaSetDevice(device), "cudaSetDevice");
    faRecords.clear();
    faStatus = -1;
    faDnaLineLen = 0;
    faMaxLastLineLen = 0;
    faLaunches = 0;
    resident.valid = false;
    if (n == 0) return;                        // kseq.h:239-240: end of file before the first header
    if (bytes[0] != '>') {                     // :239: not FASTA
        faStatus = -3;
        return;
    }
    if (n >= (1ull << 40)) throw MatcherError(MBGCG_E_UNSUPPORTED, "FASTA image of 1 TiB or more");
    if (!faEv[0])
        for (auto& e : faEv) check(cudaEventCreate(&e), "cudaEventCreate");
    faMs[0] = faMs[1] = faMs[2] = 0;

    // ---- the file image -> device: 16 bytes in front (a '\n' right before byte 0), zeros behind ----
    const uint64_t inBytes = 16 + align_up(n, 16) + 64;
    growDevice(reinterpret_cast<void**>(&dFaIn), &dFaInCap, inBytes, "cudaMalloc(fasta image)");
    uint8_t* in = dFaIn + 16;
    check(cudaEventRecord(faEv[0], stream), "event");
    check(cudaMemsetAsync(dFaIn, '\n', 16, stream), "memset");
    check(cudaMemsetAsync(in + n, 0, inBytes - 16 - n, stream), "memset");
    check(cudaMemcpyAsync(in, bytes, n, cudaMemcpyHostToDevice, stream), "H2D(fasta image)");
    check(cudaEventRecord(faEv[1], stream), "event");

    // ---- record starts ----
    const uint32_t blocks = static_cast<uint32_t>((n + FA_BLOCK_BYTES - 1) / FA_BLOCK_BYTES);
    const uint64_t cntBytes = align_up((blocks + 2ull) * 4, 256);
    growDevice(reinterpret_cast<void**>(&dFaWork), &dFaWorkCap, cntBytes, "cudaMalloc(fasta work)");
    growPinned(reinterpret_cast<void**>(&hFaWork), &hFaWorkCap, 4096, "cudaHostAlloc(fasta work)");
    uint32_t* dCnt = reinterpret_cast<uint32_t*>(dFaWork);
    int launches = launch_fa_count(in, n, dCnt, stream);
    check(cudaMemcpyAsync(hFaWork, dCnt + blocks, 4, cudaMemcpyDeviceToHost, stream), "D2H(record count)");
    check(cudaStreamSynchronize(stream), "ingest sync");
    const uint32_t nRecs = *reinterpret_cast<uint32_t*>(hFaWork);
    if (nRecs == 0) throw MatcherError(MBGCG_E_CUDA, "FASTA ingest found no record start");  // bytes[0] == '>' is one

    const uint64_t oStarts = cntBytes;
    const uint64_t oRecs = align_up(oStarts + nRecs * 8ull, 256);
    const uint64_t oGather = align_up(oRecs + nRecs * sizeof(FaRecord), 256);
    const uint64_t workBytes = oGather + nRecs * sizeof(FaGather) + 256;
    if (workBytes > dFaWorkCap) {  // grow, keeping the block offsets
        uint8_t* grown = nullptr;
        const uint64_t cap = align_up(workBytes + workBytes / 4, 1 << 20);
        check(cudaMalloc(reinterpret_cast<void**>(&grown), cap), "cudaMalloc(fasta work)");
        check(cudaMemcpyAsync(grown, dFaWork, cntBytes, cudaMemcpyDeviceToDevice, stream), "D2D(fasta work)");
        check(cudaStreamSynchronize(stream), "ingest sync");
        cudaFree(dFaWork);
        dFaWork = grown;
        dFaWorkCap = cap;
        dCnt = reinterpret_cast<uint32_t*>(dFaWork);
    }
    uint64_t* dStarts = reinterpret_cast<uint64_t*>(dFaWork + oStarts);
    FaRecord* dRecs = reinterpret_cast<FaRecord*>(dFaWork + oRecs);
    FaGather* dGather = reinterpret_cast<FaGather*>(dFaWork + oGather);
    launches += launch_fa_starts(in, n, dCnt, dStarts, stream);
    launches += launch_fa_records(in, n, dStarts, nRecs, dRecs, stream);
    const uint64_t hostBytes = nRecs * (sizeof(FaRecord) + sizeof(FaGather)) + 64;
    growPinned(reinterpret_cast<void**>(&hFaWork), &hFaWorkCap, hostBytes, "cudaHostAlloc(fasta work)");
    FaRecord* hRecs = reinterpret_cast<FaRecord*>(hFaWork);
    FaGather* hGather = reinterpret_cast<FaGather*>(hRecs + nRecs);
    check(cudaMemcpyAsync(hRecs, dRecs, nRecs * sizeof(FaRecord), cudaMemcpyDeviceToHost, stream), "D2H(records)");
    check(cudaStreamSynchronize(stream), "ingest sync");

    // ---- the reader's file-wide state, record by record (kseq.h:253-256, 260-265) ----
    uint64_t dna = 0, maxLast = 0, cursor = PAD, chunks = 0;
    uint32_t accepted = 0;
    for (uint32_t r = 0; r < nRecs; r++) {
        const FaRecord& rec = hRecs[r];
        if (rec.start + 1 >= n) break;  // '>' is the last byte: :244 returns end of file, the record does not exist
        const uint64_t seqStart = rec.hdrEnd + 1;
        const uint64_t len = rec.regionEnd > seqStart ? rec.regionEnd - seqStart : 0;
        const uint64_t body = len - (rec.endsWithNl ? 1 : 0);  // without the final line end
        const uint64_t line = rec.firstLine;
        bool bad = rec.bad != 0;
        uint64_t seqLen = 0, lastLen = 0;
        if (!bad && body) {
            const uint64_t fullLines = body / (line + 1);
            seqLen = body - fullLines;
            lastLen = body - fullLines * (line + 1);
            if (fullLines) {  // lines that are not the last one: they carry the file's line length (:253-256)
                if (dna == 0) dna = line;
                else if (dna != line) bad = true;
            }
        }
        if (lastLen > maxLast) maxLast = lastLen;          // :260-263
        if (!bad && dna && dna < maxLast) bad = true;       // :264
        if (bad) {
            dna = UINT64_MAX;                               // DNA_NOT_WELLFORMED
            faStatus = -4;
            break;
        }
        if (seqLen >= (1ull << 32)) throw MatcherError(MBGCG_E_UNSUPPORTED, "sequence of 4 GiB or more");
        hGather[accepted] = FaGather{seqStart, cursor, seqLen, seqLen ? line : 0, chunks};
        faRecords.push_back(FastaRecord{rec.start + 1, std::min(rec.hdrEnd, n) - (rec.start + 1), cursor, seqLen});
        const uint64_t next = align_up(cursor + seqLen + TEXT_GAP, 16);
        chunks += (next - cursor) / 16;  // the text and the zero gap behind it: the gather writes both
        cursor = next;
        accepted++;
    }
    faDnaLineLen = dna;
    faMaxLastLineLen = maxLast;

    // ---- sequences -> texts ----
    const uint64_t textBytes = cursor + TEXT_TAIL;
    growDevice(reinterpret_cast<void**>(&dFaText), &dFaTextCap, textBytes, "cudaMalloc(ingested texts)");
    check(cudaMemsetAsync(dFaText, 0, PAD, stream), "memset(ingested texts, head)");
    check(cudaMemsetAsync(dFaText + cursor, 0, TEXT_TAIL, stream), "memset(ingested texts, tail)");
    if (accepted) {
        check(cudaMemcpyAsync(dGather, hGather, accepted * sizeof(FaGather), cudaMemcpyHostToDevice, stream), "H2D(gather)");
        launches += launch_fa_gather(in, dGather, accepted, chunks, dFaText, stream);
    }
    check(cudaGetLastError(), "ingest launch");
    check(cudaEventRecord(faEv[2], stream), "event");
    check(cudaStreamSynchronize(stream), "ingest sync");
    check(cudaEventElapsedTime(&faMs[0], faEv[0], faEv[1]), "event time");  // image host -> device
    check(cudaEventElapsedTime(&faMs[1], faEv[1], faEv[2]), "event time");  // kernels (with the two small host round trips)
    check(cudaEventElapsedTime(&faMs[2], faEv[0], faEv[2]), "event time");
    faLaunches = static_cast<uint64_t>(launches);
    totalLaunches += faLaunches;
}

void B200SparseEMMatcher::copyIngested(uint64_t record, char* out) {
    check(cudaSetDevice(device), "cudaSetDevice");
    if (record >= faRecords.size()) throw MatcherError(MBGCG_E_PARAM, "no such record in the last ingest");
    if (faRecords[record].seqLen == 0) return;
    check(cudaMemcpyAsync(out, dFaText + faRecords[record].textOff, faRecords[record].seqLen, cudaMemcpyDeviceToHost, stream), "D2H(sequence)");
    check(cudaStreamSynchronize(stream), "sequence copy sync");
}

}  // namespace mbgc_b200
