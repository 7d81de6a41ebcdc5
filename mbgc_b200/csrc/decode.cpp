// mbgc_b200/csrc/decode.cpp -- host side of row f4: the decoder's match-copy path (/root/reference/mbgccoder/MBGC_Decoder.cpp:
// initReference :243-262, decodeReference :264-317, decodeTarget :532-649, loadRef :651-675).  The device finds where every
// target starts in the six streams, replays the control flow of all targets at once and writes the DNA (decode.cu).  What
// stays here is what is sequential over CONTIGS, not over bases: which contigs extend the reference (matching/MGMP_Params.h:
// 179-190 on the parsed lengths and unmatched counts), where loadRef puts them (lock positions, truncation, the wrap of the
// circular buffer), and therefore which contigs can be written in one launch -- a contig may copy from a contig decoded
// before it, so the copy pass runs in batches that end at every change of the reference's content.
#include "decode.hpp"

#include <algorithm>
#include <cstring>
#include <string>

#include "../../include/mbgc_b200.h"
#include "kernels.h"
#include "rcmatch.h"

namespace mbgc_b200 {

namespace {

void check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw MatcherError(MBGCG_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

constexpr uint64_t PAD = 256;  // readable bytes behind every device buffer the copy pass loads 16 bytes at a time from
constexpr uint8_t SEQ_SEPARATOR_MARK = static_cast<uint8_t>('"' + 128);   // mbgccoder/MBGC_Params.h:45-48
constexpr uint8_t MATCH_MARK = static_cast<uint8_t>('%' + 128);
constexpr uint8_t FILE_SEPARATOR_MARK = static_cast<uint8_t>(';' + 128);
constexpr uint64_t REF_SHIFT = 1;

}  // namespace

// the buffers only grow: a decoder that is called again (the next archive, the next block) allocates nothing
void DevBuf::reset(uint64_t bytes) {
    if (bytes <= cap) {
        n = bytes;
        return;
    }
    if (p) cudaFree(p);
    p = nullptr;
    n = cap = 0;
    check(cudaMalloc(&p, bytes), "cudaMalloc(decode)");
    n = cap = bytes;
}

DevBuf::~DevBuf() {
    if (p) cudaFree(p);
}

Decoder::Decoder(B200SparseEMMatcher* matcher) : m(matcher) {}

Decoder::~Decoder() {}

uint8_t* Decoder::upload(DevBuf& b, const uint8_t* src, uint64_t n, cudaStream_t st) {
    b.reset(n + PAD);
    check(cudaMemsetAsync(b.as<uint8_t>() + n, 0, PAD, st), "memset(stream pad)");
    if (n) check(cudaMemcpyAsync(b.p, src, n, cudaMemcpyHostToDevice, st), "H2D(stream)");
    return b.as<uint8_t>();
}

// exclusive scan of cnt[0 .. n) into out[0 .. n] (out[n] = total)
int Decoder::scan(const uint32_t* cnt, uint32_t* out, uint32_t n, cudaStream_t st) {
    check(cudaMemsetAsync(dScanState.p, 0, dScanState.n, st), "memset(scan state)");
    return launch_scan_u32(cnt, out, n, dScanState.as<unsigned long long>(), st);
}

void Decoder::run(const DcParams& p, const DecodeStreamsHost& hs, const DecodeLayout& L) {
    check(cudaSetDevice(m->deviceIndex()), "cudaSetDevice");
    cudaStream_t st = static_cast<cudaStream_t>(m->streamHandle());
    launches = 0;
    ms[0] = ms[1] = ms[2] = 0;
    contigLen.clear();
    contigUnmatched.clear();
    contigBase.clear();
    dnaLen = 0;
    for (int i = 0; i < 6; i++)
        if (hs.n[i] >= (1ull << 32)) throw MatcherError(MBGCG_E_UNSUPPORTED, "decode: a stream of 4 GiB or more (32-bit stream positions)");
    if (hs.n[1] % 4) throw MatcherError(MBGCG_E_PARAM, "decode: the offset stream is not a whole number of 32-bit entries");
    if (L.refTotalLength < REF_SHIFT) throw MatcherError(MBGCG_E_PARAM, "decode: refTotalLength");
    const bool hasG0 = hs.n[0] != 0;  // initReference / decodeReference on an empty literal stream seed nothing (:249)
    const uint32_t nG0 = !hasG0 ? 0 : (L.g0IsTarget ? 1 : L.g0SeqCount);
    uint64_t nContigs64 = nG0;
    for (uint32_t t = 0; t < L.nTargets; t++) nContigs64 += L.seqCounts[t];
    if (nContigs64 >= (1ull << 28)) throw MatcherError(MBGCG_E_UNSUPPORTED, "decode: 2^28 contigs or more");
    const uint32_t nContigs = static_cast<uint32_t>(nContigs64);

    cudaEvent_t ev[4];
    for (auto& e : ev) check(cudaEventCreate(&e), "cudaEventCreate");
    struct EvGuard {
        cudaEvent_t* e;
        ~EvGuard() { for (int i = 0; i < 4; i++) cudaEventDestroy(e[i]); }
    } evGuard{ev};
    check(cudaEventRecord(ev[0], st), "event");
    dc_upload_tables(st);

    // ---- streams to the device ----
    DcStreams s{};
    s.lit = upload(dLit, hs.p[0], hs.n[0], st);
    s.nLit = hs.n[0];
    s.off = upload(dOff, hs.p[1], hs.n[1], st);
    s.nOff = hs.n[1] / 4;
    s.off5 = upload(dOff5, hs.p[2], hs.n[2], st);
    s.nOff5 = hs.n[2];
    const uint8_t* dLenRaw = upload(dLenSrc, hs.p[3], hs.n[3], st);
    s.gd = upload(dGd, hs.p[4], hs.n[4], st);
    s.nGd = hs.n[4];
    s.flags = upload(dFlags, hs.p[5], hs.n[5], st);
    s.nFlags = hs.n[5];

    // ---- where the contigs and targets start ----
    const uint32_t litBlocks = static_cast<uint32_t>((hs.n[0] + DC_COUNT_BLOCK - 1) / DC_COUNT_BLOCK);
    const uint32_t flagBlocks = static_cast<uint32_t>((hs.n[5] + DC_COUNT_BLOCK - 1) / DC_COUNT_BLOCK);
    const uint32_t maxBlocks = std::max({litBlocks, flagBlocks, nContigs, 1u});
    dScanState.reset((static_cast<uint64_t>(maxBlocks) / 4096 + 16) * 8);
    dCnt0.reset((static_cast<uint64_t>(maxBlocks) + 2) * 4);
    dCnt1.reset((static_cast<uint64_t>(maxBlocks) + 2) * 4);
    dOff0.reset((static_cast<uint64_t>(maxBlocks) + 2) * 4);
    dOff1.reset((static_cast<uint64_t>(maxBlocks) + 2) * 4);
    uint32_t nSep = 0, nMarks = 0, nFlagSep = 0;
    if (litBlocks) {
        launches += launch_dc_count(s.lit, s.nLit, SEQ_SEPARATOR_MARK, MATCH_MARK, dCnt0.as<uint32_t>(), dCnt1.as<uint32_t>(), st);
        launches += scan(dCnt0.as<uint32_t>(), dOff0.as<uint32_t>(), litBlocks, st);
        launches += scan(dCnt1.as<uint32_t>(), dOff1.as<uint32_t>(), litBlocks, st);
        check(cudaMemcpyAsync(&nSep, dOff0.as<uint32_t>() + litBlocks, 4, cudaMemcpyDeviceToHost, st), "D2H");
        check(cudaMemcpyAsync(&nMarks, dOff1.as<uint32_t>() + litBlocks, 4, cudaMemcpyDeviceToHost, st), "D2H");
        check(cudaStreamSynchronize(st), "sync(marker counts)");
    }
    if (nSep < nContigs)
        throw MatcherError(MBGCG_E_PARAM, "decode: the literal stream holds fewer sequences than the layout names");
    dSepPos.reset((static_cast<uint64_t>(nSep) + 2) * 4);
    dSepMarks.reset((static_cast<uint64_t>(nSep) + 2) * 4);
    dMarkPos.reset((static_cast<uint64_t>(nMarks) + 2) * 4);
    if (litBlocks)
        launches += launch_dc_scatter(s.lit, s.nLit, SEQ_SEPARATOR_MARK, MATCH_MARK, dOff0.as<uint32_t>(), dOff1.as<uint32_t>(),
                                      dSepPos.as<uint32_t>(), dSepMarks.as<uint32_t>(), dMarkPos.as<uint32_t>(), st);
    if (flagBlocks) {
        launches += launch_dc_count(s.flags, s.nFlags, FILE_SEPARATOR_MARK, 0x100, dCnt0.as<uint32_t>(), nullptr, st);
        launches += scan(dCnt0.as<uint32_t>(), dOff0.as<uint32_t>(), flagBlocks, st);
        check(cudaMemcpyAsync(&nFlagSep, dOff0.as<uint32_t>() + flagBlocks, 4, cudaMemcpyDeviceToHost, st), "D2H");
        check(cudaStreamSynchronize(st), "sync(flag separators)");
    }
    dFlagSep.reset((static_cast<uint64_t>(nFlagSep) + 2) * 4);
    if (flagBlocks)
        launches += launch_dc_scatter(s.flags, s.nFlags, FILE_SEPARATOR_MARK, 0x100, dOff0.as<uint32_t>(), nullptr, dFlagSep.as<uint32_t>(),
                                      nullptr, nullptr, st);
    if (p.enableExt && nFlagSep < L.nTargets)
        throw MatcherError(MBGCG_E_PARAM, "decode: the flag stream holds fewer file separators than the layout names targets");
    // gap deltas and omitted offsets in front of every contig
    dGdStart.reset((static_cast<uint64_t>(nContigs) + 2) * 4);
    dNzStart.reset((static_cast<uint64_t>(nContigs) + 2) * 4);
    if (nContigs) {
        launches += launch_dc_contig_gd(nContigs, dSepMarks.as<uint32_t>(), p.gapDepthOffset, dCnt0.as<uint32_t>(), st);
        launches += scan(dCnt0.as<uint32_t>(), dGdStart.as<uint32_t>(), nContigs, st);
        launches += launch_dc_contig_nz(nContigs, dGdStart.as<uint32_t>(), s.gd, s.nGd, dCnt1.as<uint32_t>(), st);
        launches += scan(dCnt1.as<uint32_t>(), dNzStart.as<uint32_t>(), nContigs, st);
    }
    // the match lengths as 32-bit values (decodeMapLenStream :966-986)
    if (L.frugalLen) {
        const uint64_t cap = hs.n[3] / 2 + 1;
        dLen.reset(cap * 4 + PAD);
        dLenChunk.reset((dc_len_chunks(hs.n[3]) + 1) * 32);
        dLenBlock.reset((dc_len_blocks(hs.n[3]) + 1) * 32);
        dLenEntry.reset((dc_len_blocks(hs.n[3]) + 2) * 8);
        uint64_t* dTotal = dLenEntry.as<uint64_t>() + dc_len_blocks(hs.n[3]) + 1;
        launches += launch_dc_len_expand(dLenRaw, hs.n[3], dLenChunk.as<uint32_t>(), dLenBlock.as<uint32_t>(), dLenEntry.as<uint64_t>(), dTotal,
                                         dLen.as<uint32_t>(), cap, st);
        uint64_t total = 0;
        check(cudaMemcpyAsync(&total, dTotal, 8, cudaMemcpyDeviceToHost, st), "D2H");
        check(cudaStreamSynchronize(st), "sync(length stream)");
        s.len = dLen.as<uint32_t>();
        s.nLen = total;
    } else {
        s.len = reinterpret_cast<const uint32_t*>(dLenRaw);
        s.nLen = hs.n[3] / 4;
    }

    // ---- the parse: one thread per target (and one for the G0 prologue) ----
    std::vector<DcTarget> targets;
    uint32_t c = 0;
    if (nG0) {
        targets.push_back(DcTarget{0, nG0, UINT32_MAX, 0, L.refTotalLength});
        c = nG0;
    }
    for (uint32_t t = 0; t < L.nTargets; t++) {
        targets.push_back(DcTarget{c, L.seqCounts[t], p.enableExt ? t : UINT32_MAX, 0, L.lockPos ? L.lockPos[t] : L.refTotalLength});  // no extensions: no flag stream
        c += L.seqCounts[t];
    }
    dTargets.reset(targets.size() * sizeof(DcTarget) + 16);
    if (!targets.empty())
        check(cudaMemcpyAsync(dTargets.p, targets.data(), targets.size() * sizeof(DcTarget), cudaMemcpyHostToDevice, st), "H2D(targets)");
    const uint64_t nSegs = 4ull * nMarks + nContigs;
    if (nSegs >= (1ull << 32)) throw MatcherError(MBGCG_E_UNSUPPORTED, "decode: 2^30 matches or more");
    dSegs.reset((nSegs + 1) * sizeof(DcSeg));
    check(cudaMemsetAsync(dSegs.p, 0, dSegs.n, st), "memset(segments)");
    dContig.reset((static_cast<uint64_t>(nContigs) + 1) * 16);
    check(cudaMemsetAsync(dContig.p, 0, dContig.n, st), "memset(contigs)");
    dErr.reset(16);
    check(cudaMemsetAsync(dErr.p, 0, 16, st), "memset(error)");
    DcIndex ix{dSepPos.as<uint32_t>(), dSepMarks.as<uint32_t>(), dMarkPos.as<uint32_t>(), dGdStart.as<uint32_t>(), dNzStart.as<uint32_t>(),
               dFlagSep.as<uint32_t>(), nContigs, nFlagSep};
    uint64_t* dLenOut = dContig.as<uint64_t>();
    uint64_t* dUnmOut = dContig.as<uint64_t>() + nContigs;
    launches += launch_dc_parse(p, s, ix, dTargets.as<DcTarget>(), static_cast<uint32_t>(targets.size()), dSegs.as<DcSeg>(), dLenOut, dUnmOut,
                                dErr.as<uint32_t>(), st);
    contigLen.resize(nContigs);
    contigUnmatched.resize(nContigs);
    std::vector<uint32_t> sepMarks(nContigs);
    uint32_t err = 0;
    if (nContigs) {
        check(cudaMemcpyAsync(contigLen.data(), dLenOut, 8ull * nContigs, cudaMemcpyDeviceToHost, st), "D2H(contig lengths)");
        check(cudaMemcpyAsync(contigUnmatched.data(), dUnmOut, 8ull * nContigs, cudaMemcpyDeviceToHost, st), "D2H(unmatched)");
        check(cudaMemcpyAsync(sepMarks.data(), dSepMarks.p, 4ull * nContigs, cudaMemcpyDeviceToHost, st), "D2H(marks)");
    }
    check(cudaMemcpyAsync(&err, dErr.p, 4, cudaMemcpyDeviceToHost, st), "D2H(error)");
    check(cudaEventRecord(ev[1], st), "event");
    check(cudaStreamSynchronize(st), "sync(parse)");
    for (uint32_t i = 0; i < nG0; i++) contigUnmatched[i] = 0;  // verbatim contigs: no decodeSequence call counted them
    if (err) throw MatcherError(MBGCG_E_PARAM, "decode: the streams are inconsistent (a cursor ran off its stream or off the reference)");

    // ---- the copy pass, in batches that end where the reference changes ----
    contigBase.assign(static_cast<size_t>(nContigs) + 1, 0);
    for (uint32_t i = 0; i < nContigs; i++) contigBase[i + 1] = contigBase[i] + contigLen[i];
    dnaLen = contigBase[nContigs];
    firstOutputContig = L.g0IsTarget ? nG0 : 0;
    dDna.reset(dnaLen + PAD);
    dRef.reset(L.refTotalLength + 2 * PAD);
    check(cudaMemsetAsync(dRef.p, 0, dRef.n, st), "memset(reference)");
    dBase.reset((static_cast<uint64_t>(nContigs) + 1) * 8);
    check(cudaMemcpyAsync(dBase.p, contigBase.data(), (static_cast<uint64_t>(nContigs) + 1) * 8, cudaMemcpyHostToDevice, st), "H2D(bases)");
    dTileSeg.reset(((dnaLen >> DC_TILE_SHIFT_MIN) + 2) * 4);
    uint8_t* ref = dRef.as<uint8_t>();
    uint8_t* dna = dDna.as<uint8_t>();
    refPos = REF_SHIFT;
    reachedCount = 0;
    batches = 0;
    uint32_t batchStart = 0;
    auto marksBefore = [&](uint32_t i) -> uint64_t { return i ? sepMarks[i - 1] : 0; };
    auto flush = [&](uint32_t cEnd) {
        if (cEnd <= batchStart) return;
        const uint64_t segBegin = 4 * marksBefore(batchStart) + batchStart, segEnd = 4 * marksBefore(cEnd) + cEnd;
        const uint64_t d0 = contigBase[batchStart], d1 = contigBase[cEnd];
        if (d1 > d0) {
            // long matches: large tiles (the walk over a tile's segments is overhead per tile); short ones: small tiles (the
            // walk is the work, and it is serial within a tile)
            const uint64_t matches = (segEnd - segBegin - (cEnd - batchStart)) / 4;
            const uint64_t mean = (d1 - d0) / std::max<uint64_t>(matches, 1);
            const int shift = mean >= 2048 ? DC_TILE_SHIFT_MAX : (mean >= 256 ? 14 : DC_TILE_SHIFT_MIN);
            launches += launch_dc_tiles(dSegs.as<DcSeg>(), segBegin, segEnd, dBase.as<uint64_t>(), d0, shift, dTileSeg.as<uint32_t>(), st);
            launches += launch_dc_copy(dSegs.as<DcSeg>(), segEnd, dBase.as<uint64_t>(), dTileSeg.as<uint32_t>(), shift, d0, d1, ref, s, dna, st);
            batches++;
        }
        batchStart = cEnd;
    };
    // loadRef (:651-675): `text` = device pointer of the source (a decoded contig, a stretch of the reference, or nullptr for
    // the one-byte region separator); upTo = the contig that must be on the device before the load can read it
    auto loadRef = [&](const uint8_t* text, uint64_t seqLength, uint64_t refLockPos, bool rc, uint32_t upTo) {
        uint64_t consumed = 0;
        while (seqLength != 0) {
            if (refPos == L.refTotalLength && refLockPos != L.refTotalLength) {
                reachedCount++;
                refPos = REF_SHIFT;
            }
            uint64_t tmpLength = seqLength;
            const uint64_t tmpMax = refLockPos < refPos ? L.refTotalLength : refLockPos;
            if (refPos + tmpLength > tmpMax) tmpLength = tmpMax - refPos;
            if (tmpLength) {
                if (!text) {
                    // REF_REGION_SEPARATOR = 0 on memory that is still zero: nothing to write, nothing to wait for
                    if (reachedCount) {
                        flush(upTo);
                        check(cudaMemsetAsync(ref + refPos, 0, tmpLength, st), "memset(separator)");
                    }
                } else {
                    flush(upTo);
                    if (rc)
                        launch_ref_segment(text + seqLength - tmpLength, ref + refPos, tmpLength, true, st);
                    else
                        check(cudaMemcpyAsync(ref + refPos, text + consumed, tmpLength, cudaMemcpyDeviceToDevice, st), "D2D(reference)");
                    launches++;
                }
            }
            refPos += tmpLength;
            if (!rc) consumed += tmpLength;
            seqLength = refPos == refLockPos ? 0 : seqLength - tmpLength;
            if (p.lazy && refPos == refLockPos) {
                flush(upTo);
                check(cudaMemsetAsync(ref + refLockPos - 1, 0, 1, st), "memset(separator)");
            }
        }
    };
    if (nG0) {
        // initReference / decodeReference: the G0 contigs, verbatim, and the reverse complement of all of them
        flush(nG0);
        const uint64_t g0 = contigBase[nG0];
        if (REF_SHIFT + g0 * (L.rcInReference ? 2 : 1) > L.refTotalLength)
            throw MatcherError(MBGCG_E_PARAM, "decode: G0 does not fit the reference");
        if (g0) {
            check(cudaMemcpyAsync(ref + REF_SHIFT, dna, g0, cudaMemcpyDeviceToDevice, st), "D2D(G0)");
            launches++;
            if (L.rcInReference) {
                launch_ref_segment(dna, ref + REF_SHIFT + g0, g0, true, st);
                launches++;
            }
        }
        refPos = REF_SHIFT + g0 * (L.rcInReference ? 2 : 1);
    }
    c = nG0;
    for (uint32_t t = 0; t < L.nTargets; t++) {
        const uint64_t lock = L.lockPos ? L.lockPos[t] : L.refTotalLength;
        const uint64_t startPos = refPos;
        const uint64_t factor = L.factors[2 * t], rcFactor = L.factors[2 * t + 1];
        for (uint32_t k = 0; k < L.seqCounts[t]; k++, c++) {
            const uint64_t n = contigLen[c], u = contigUnmatched[c];
            const bool load = u * factor > n;  // isContigProperForRefExtension, MGMP_Params.h:179-186
            if (load) loadRef(dna + contigBase[c], n, lock, false, c + 1);
            if (L.rcInReference && L.contigsIndividuallyReversed && u * rcFactor > n && load) loadRef(dna + contigBase[c], n, lock, true, c + 1);
        }
        if (L.rcInReference && !L.contigsIndividuallyReversed) {  // :627-635
            if (refPos >= startPos) {
                loadRef(ref + startPos, refPos - startPos, lock, true, c);
            } else {
                const uint64_t tmpStartPos = startPos;
                loadRef(ref + 1, refPos - 1, lock, true, c);
                loadRef(ref + tmpStartPos, L.refTotalLength - tmpStartPos, lock, true, c);
            }
        }
        if (p.lazy) loadRef(nullptr, 1, lock, false, c);
    }
    flush(nContigs);
    check(cudaEventRecord(ev[2], st), "event");
    check(cudaStreamSynchronize(st), "sync(copy pass)");
    check(cudaEventElapsedTime(&ms[0], ev[0], ev[1]), "elapsed");
    check(cudaEventElapsedTime(&ms[1], ev[1], ev[2]), "elapsed");
    refTotal = L.refTotalLength;
}

void Decoder::fetch(uint32_t firstContig, uint32_t n, uint8_t* dst) {
    if (static_cast<uint64_t>(firstContig) + n > contigLen.size()) throw MatcherError(MBGCG_E_PARAM, "decode_fetch: contig range");
    check(cudaSetDevice(m->deviceIndex()), "cudaSetDevice");
    cudaStream_t st = static_cast<cudaStream_t>(m->streamHandle());
    const uint64_t a = contigBase[firstContig], b = contigBase[firstContig + n];
    if (b > a) check(cudaMemcpyAsync(dst, dDna.as<uint8_t>() + a, b - a, cudaMemcpyDeviceToHost, st), "D2H(dna)");
    check(cudaStreamSynchronize(st), "sync(fetch)");
}

void Decoder::fetchRef(uint64_t pos, uint64_t n, uint8_t* dst) {
    if (pos + n > refTotal) throw MatcherError(MBGCG_E_PARAM, "decode_fetch_ref: range");
    check(cudaSetDevice(m->deviceIndex()), "cudaSetDevice");
    cudaStream_t st = static_cast<cudaStream_t>(m->streamHandle());
    if (n) check(cudaMemcpyAsync(dst, dRef.as<uint8_t>() + pos, n, cudaMemcpyDeviceToHost, st), "D2H(reference)");
    check(cudaStreamSynchronize(st), "sync(fetch)");
}

}  // namespace mbgc_b200
