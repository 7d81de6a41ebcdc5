// mbgc_b200/csrc/packer.cpp -- host side of K6: B200SparseEMMatcher::packBatch (see matcher.hpp, pack.cu).
//
// Mirrors MBGC_Encoder::processMatches (/root/reference/mbgccoder/MBGC_Encoder.cpp:143-308) for n texts per call:
// same inputs (match list, text, matchingLockPos, refExtLoadedPosArr, the matcher's reference / cursors), same
// outputs (the bytes appended to targetLiterals, targetMapOffDests, targetMapOff5thByte, targetMapLenDests,
// targetGapDeltas, targetGapMismatchesFlags, the surviving match list and the returned unmatched count).  What
// stays with the caller, because it needs encoder state the matcher does not own: the dissimilar-contig bail-out
// (:202-205) and the statistics counters (:289-302).  processLiteral (:128-135, a developer-build option, off by
// default) comes back as a list of text ranges when setLiteralRefExt switched it on.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "../../include/mbgc_b200.h"
#include "matcher.hpp"

namespace mbgc_b200 {

namespace {
constexpr uint64_t PAD = 64, TEXT_GAP = 64, TEXT_TAIL = 8192;  // = matcher.cpp
inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }
}  // namespace

void B200SparseEMMatcher::releasePack() {
    cudaFree(dPackWork);
    cudaFree(dPackIn);
    if (hPackIn) cudaFreeHost(hPackIn);
    if (hPackSum) cudaFreeHost(hPackSum);
    for (auto& p : hPackOut)
        if (p) cudaFreeHost(p);
    for (auto& e : pkEv)
        if (e) cudaEventDestroy(e);
}

void B200SparseEMMatcher::packBatch(const PackParams& p, const std::vector<TextView>& texts, const TextMatch* matches,
                                    const uint64_t* counts, const uint64_t* lockPos, const uint64_t* loadedPosArr,
                                    uint64_t nLoadedPos) {
    check(cudaSetDevice(device), "cudaSetDevice");
    if (p.gapDepthOffset < 0 || p.gapDepthOffset > 127)
        throw MatcherError(MBGCG_E_UNSUPPORTED, "gapDepthOffsetEncoding above 127 is outside the device packer's scope");
    if (p.enableExt && (p.gapDepthMismatches < 0 || p.gapDepthMismatches > 2))
        throw MatcherError(MBGCG_E_UNSUPPORTED, "gapDepthMismatchesEncoding above 2 is outside the device packer's scope "
                                                "(a developer-build option of the reference; the default is 2)");
    const bool useResident = matches == nullptr;
    if (useResident) {
        if (!resident.valid || resident.lens.size() != texts.size())
            throw MatcherError(MBGCG_E_PARAM, "packBatch without match lists needs the immediately preceding single-pass "
                                              "match call of the same texts");
        for (size_t i = 0; i < texts.size(); i++)
            if (texts[i].len != resident.lens[i]) throw MatcherError(MBGCG_E_PARAM, "packBatch: texts differ from the match call");
    }
    for (const auto& t : texts)
        if (t.len >= (1ull << 32)) throw MatcherError(MBGCG_E_UNSUPPORTED, "target text of 4 GiB or more");
    if (!pkEv[0])
        for (auto& e : pkEv) check(cudaEventCreate(&e), "cudaEventCreate");

    const size_t nT = texts.size();
    for (int k = 0; k < PACK_STREAMS; k++) packOff[k].assign(nT + 1, 0);
    packUnmatchedChars.assign(nT, 0);
    packKeptCounts.assign(nT, 0);
    packKeptHost.clear();
    packKeptLast = KeptSlice{nullptr, 0};
    packLitRanges.clear();
    const bool wantLit = litExt.minLen != UINT64_MAX;
    for (auto& m : packMs) m = 0;
    packLaunches = 0;
    packSweeps[0] = packSweeps[1] = 0;
    if (nT == 0) return;

    uint64_t maxPassBytes = 1024ull << 20;
    if (const char* e = std::getenv("MBGC_B200_MAX_PASS_BYTES")) maxPassBytes = std::max<uint64_t>(1, std::strtoull(e, nullptr, 10));
    uint64_t outUsed[PACK_STREAMS] = {0, 0, 0, 0, 0, 0};
    size_t first = 0;
    uint64_t matchBase = 0;
    while (first < nT) {
        size_t last = first;
        uint64_t bytes = 0;
        while (last < nT && (useResident || last == first || bytes + texts[last].len <= maxPassBytes)) bytes += texts[last++].len;
        const int nC = static_cast<int>(last - first);
        uint64_t nM64 = 0;
        if (useResident)
            nM64 = resident.nMatches;
        else
            for (int c = 0; c < nC; c++) nM64 += counts[first + c];
        if (nM64 + static_cast<uint64_t>(nC) >= 0xFFFFFF00ull) throw MatcherError(MBGCG_E_UNSUPPORTED, "more than 2^32 matches in one pass");
        const uint32_t nM = static_cast<uint32_t>(nM64);

        // ---- inputs ----
        check(cudaEventRecord(pkEv[0], stream), "event");
        PackBatch b{};
        b.nContigs = nC;
        b.nMatches = nM;
        uint64_t textBytes = 0;
        const uint64_t inBytes = align_up(static_cast<uint64_t>(nM) * sizeof(TextMatch), 16) + align_up((nC + 1) * 4ull, 16) +
                                 static_cast<uint64_t>(nC) * 24 + 16 + nLoadedPos * 8 + 64;
        growPinned(reinterpret_cast<void**>(&hPackIn), &hPackInCap, inBytes, "cudaHostAlloc(pack input)");
        growDevice(reinterpret_cast<void**>(&dPackIn), &dPackInCap, inBytes, "cudaMalloc(pack input)");
        uint64_t inOff = 0;
        auto place = [&](uint64_t n) { const uint64_t o = inOff; inOff = align_up(inOff + n, 16); return o; };
        const uint64_t oLoaded = place(nLoadedPos * 8);
        if (nLoadedPos) std::memcpy(hPackIn + oLoaded, loadedPosArr, nLoadedPos * 8);
        if (useResident) {
            b.text = resident.bd.text;
            b.contigOff = resident.bd.contigOff;
            b.contigLock = resident.bd.contigLock;
            b.contigLen = resident.bd.contigLen;
            b.matches = wb.matchesDense;
            b.matchOff = wb.matchOff;
            for (int c = 0; c < nC; c++) textBytes += texts[first + c].len;
            if (lockPos) {  // the caller's lock positions win over those of the match call
                const uint64_t oLock = place(nC * 8ull);
                std::memcpy(hPackIn + oLock, lockPos, nC * 8ull);
                b.contigLock = reinterpret_cast<const uint64_t*>(dPackIn + oLock);
            }
        } else {
            resident.valid = false;  // the device text buffer is about to be overwritten
            const uint64_t oM = place(static_cast<uint64_t>(nM) * sizeof(TextMatch));
            const uint64_t oMOff = place((nC + 1) * 4ull);
            const uint64_t oCOff = place(nC * 8ull);
            const uint64_t oLock = place(nC * 8ull);
            const uint64_t oLen = place(nC * 4ull);
            if (nM) std::memcpy(hPackIn + oM, matches + matchBase, static_cast<uint64_t>(nM) * sizeof(TextMatch));
            uint32_t* hMOff = reinterpret_cast<uint32_t*>(hPackIn + oMOff);
            uint64_t* hCOff = reinterpret_cast<uint64_t*>(hPackIn + oCOff);
            uint64_t* hLock = reinterpret_cast<uint64_t*>(hPackIn + oLock);
            uint32_t* hLen = reinterpret_cast<uint32_t*>(hPackIn + oLen);
            std::vector<uint64_t> off(nC);
            uint64_t cursor = PAD;
            uint32_t acc = 0;
            for (int c = 0; c < nC; c++) {
                off[c] = cursor;
                cursor = align_up(cursor + texts[first + c].len + TEXT_GAP, 16);
                hMOff[c] = acc;
                acc += static_cast<uint32_t>(counts[first + c]);
                hCOff[c] = off[c];
                hLock[c] = lockPos ? lockPos[first + c] : NO_LOCK;
                hLen[c] = static_cast<uint32_t>(texts[first + c].len);
            }
            hMOff[nC] = acc;
            textBytes = cursor + TEXT_TAIL;
            if (textBytes > dTextCap) {
                cudaFree(dText);
                dText = nullptr;
                dTextCap = 0;
                const uint64_t cap = align_up(textBytes + textBytes / 4, 1 << 20);
                check(cudaMalloc(&dText, cap), "cudaMalloc(text)");
                dTextCap = cap;
            }
            const int xl = stageTexts(texts, first, nC, off);
            totalLaunches += static_cast<uint64_t>(xl);
            packLaunches += static_cast<uint64_t>(xl);
            lastOffsets = off;  // mbgcg_copy_staged_text describes the texts that are on the device NOW
            lastLens.resize(nC);
            for (int c = 0; c < nC; c++) lastLens[c] = texts[first + c].len;
            b.text = dText;
            b.matches = reinterpret_cast<const TextMatchRec*>(dPackIn + oM);
            b.matchOff = reinterpret_cast<const uint32_t*>(dPackIn + oMOff);
            b.contigOff = reinterpret_cast<const uint64_t*>(dPackIn + oCOff);
            b.contigLock = lockPos ? reinterpret_cast<const uint64_t*>(dPackIn + oLock) : nullptr;
            b.contigLen = reinterpret_cast<const uint32_t*>(dPackIn + oLen);
        }
        if (inOff) check(cudaMemcpyAsync(dPackIn, hPackIn, inOff, cudaMemcpyHostToDevice, stream), "H2D(pack input)");
        PackState s{};
        s.ref = dRef;
        s.loadingPos = getLoadingPosition();
        s.maxRefLength = maxRefLength;
        s.refLength = getRefLength();
        s.loadedArr = nLoadedPos ? reinterpret_cast<const uint64_t*>(dPackIn + oLoaded) : nullptr;
        s.nLoaded = static_cast<uint32_t>(nLoadedPos);

        // ---- work memory ----
        const uint64_t nGaps = static_cast<uint64_t>(nM) + nC;
        const uint64_t blocks = (nGaps + PK_THREADS - 1) / PK_THREADS;
        const uint32_t longCap = static_cast<uint32_t>(textBytes / PK_LONG_RUN + nC + 16);
        uint64_t woff = 0;
        auto carve = [&](uint64_t n) { const uint64_t o = woff; woff = align_up(woff + n, 256); return o; };
        const uint64_t oFlags = carve(64 * 4);
        const uint64_t oSumLen = carve(nC * 8ull);
        const uint64_t oNl = carve(nM * 8ull);
        const uint64_t oLitScan = carve(wantLit ? (nGaps / 4096 + 8) * 8 : 0);
        const uint64_t oZeroEnd = woff;  // [0, oZeroEnd) is zeroed before every run
        const uint64_t oMContig = carve(nM * 4ull);
        const uint64_t oRemoved = carve(nM);
        const uint64_t oKeptScan = carve((nM + 1ull) * 4);
        const uint64_t oKept = carve(nM * sizeof(KeptRec));
        const uint64_t oKContig = carve(nM * 4ull);
        const uint64_t oKDiag = carve(nM * 4ull);
        const uint64_t oKOff = carve((nC + 1ull) * 4);
        const uint64_t oClaim = carve(nM);
        const uint64_t oDelta = carve(nM);
        const uint64_t oClaimedBy = carve(nM * 4ull);
        const uint64_t oGapCnt = carve(nGaps * PACK_STREAMS * 4);
        const uint64_t oBlockTot = carve((blocks + 2) * PACK_STREAMS * 4);
        const uint64_t oStreamOff = carve(static_cast<uint64_t>(PACK_STREAMS) * (nC + 1) * 4);
        const uint64_t oLong = carve(static_cast<uint64_t>(longCap) * sizeof(LongRun));
        const uint64_t oLitSpan = carve(wantLit ? nGaps * sizeof(uint2) : 0);
        const uint64_t oLitFlag = carve(wantLit ? (nGaps + 1) * 4 : 0);
        const uint64_t oLitRanges = carve(wantLit ? nGaps * sizeof(LitRange) : 0);
        uint64_t cap[PACK_STREAMS];
        uint64_t sumText = 0;
        for (int c = 0; c < nC; c++) sumText += texts[first + c].len;
        cap[PK_LITERALS] = sumText + nM + 64;
        cap[PK_MAPOFF] = nM * 4ull + 64;
        cap[PK_MAPOFF5] = nM + 64ull;
        cap[PK_MAPLEN] = nM * 14ull + 64;
        cap[PK_GAPDELTA] = nM + 64ull;
        cap[PK_GAPFLAGS] = sumText + 2 * nGaps + 64;
        uint64_t oOut[PACK_STREAMS];
        for (int k = 0; k < PACK_STREAMS; k++) oOut[k] = carve(cap[k]);
        growDevice(reinterpret_cast<void**>(&dPackWork), &dPackWorkCap, woff, "cudaMalloc(pack work)");
        PackWork w{};
        w.flags = reinterpret_cast<uint32_t*>(dPackWork + oFlags);
        w.sumLen = reinterpret_cast<unsigned long long*>(dPackWork + oSumLen);
        w.nl = reinterpret_cast<unsigned long long*>(dPackWork + oNl);
        w.mContig = reinterpret_cast<uint32_t*>(dPackWork + oMContig);
        w.removed = dPackWork + oRemoved;
        w.keptScan = reinterpret_cast<uint32_t*>(dPackWork + oKeptScan);
        w.kept = reinterpret_cast<KeptRec*>(dPackWork + oKept);
        w.kContig = reinterpret_cast<uint32_t*>(dPackWork + oKContig);
        w.kDiag = reinterpret_cast<uint32_t*>(dPackWork + oKDiag);
        w.kOff = reinterpret_cast<uint32_t*>(dPackWork + oKOff);
        w.claim = dPackWork + oClaim;
        w.delta = dPackWork + oDelta;
        w.claimedBy = reinterpret_cast<uint32_t*>(dPackWork + oClaimedBy);
        w.gapCnt = reinterpret_cast<uint32_t*>(dPackWork + oGapCnt);
        w.blockTot = reinterpret_cast<uint32_t*>(dPackWork + oBlockTot);
        w.streamOff = reinterpret_cast<uint32_t*>(dPackWork + oStreamOff);
        w.longRuns = reinterpret_cast<LongRun*>(dPackWork + oLong);
        w.longCap = longCap;
        for (int k = 0; k < PACK_STREAMS; k++) w.out[k] = dPackWork + oOut[k];
        if (wantLit) {
            w.litSpan = reinterpret_cast<uint2*>(dPackWork + oLitSpan);
            w.litFlag = reinterpret_cast<uint32_t*>(dPackWork + oLitFlag);
            w.litRanges = reinterpret_cast<LitRange*>(dPackWork + oLitRanges);
            w.litScanState = reinterpret_cast<unsigned long long*>(dPackWork + oLitScan);
        }

        // summary that comes back before the streams: flags, stream totals, kept offsets, kept length sums, offsets
        const uint64_t sumBytes = 64 * 4 + PACK_STREAMS * 4 + 8 + (nC + 1ull) * 4 + 8 + nC * 8ull +
                                  static_cast<uint64_t>(PACK_STREAMS) * (nC + 1) * 4 + 64;
        growPinned(reinterpret_cast<void**>(&hPackSum), &hPackSumCap, sumBytes, "cudaHostAlloc(pack summary)");
        uint32_t* hFlags = reinterpret_cast<uint32_t*>(hPackSum);
        uint32_t* hTot = hFlags + 64;
        unsigned long long* hSumLen = reinterpret_cast<unsigned long long*>(hPackSum + align_up(64 * 4 + PACK_STREAMS * 4, 8));
        uint32_t* hKOff = reinterpret_cast<uint32_t*>(hSumLen + nC);
        uint32_t* hStreamOff = hKOff + (nC + 1) + ((nC + 1) & 1);
        uint32_t* hLitCount = hStreamOff + static_cast<uint64_t>(PACK_STREAMS) * (nC + 1);

        check(cudaEventRecord(pkEv[1], stream), "event");
        // two blind sweeps each: the second only confirms the first.  MBGC_B200_PACK_SWEEPS=1 starts with one (whose
        // flag is then always set) and so exercises the re-run path below (tests).
        int remSweeps = 2, claimSweeps = 2;
        if (const char* e = std::getenv("MBGC_B200_PACK_SWEEPS")) remSweeps = claimSweeps = std::max(1, std::atoi(e));
        for (int attempt = 0;; attempt++) {
            check(cudaMemsetAsync(dPackWork, 0, oZeroEnd, stream), "memset(pack state)");
            if (nM == 0) check(cudaMemsetAsync(w.kOff, 0, (nC + 1ull) * 4, stream), "memset(kOff)");
            int launches = 0;
            launches += launch_pack_pass1(p, s, b, w, remSweeps, stream);
            launches += launch_pack_claims(p, s, b, w, claimSweeps, stream);
            launches += launch_pack_count(p, s, b, w, stream);
            launches += launch_pack_emit(p, s, b, w, stream);
            if (wantLit) launches += launch_pack_literal_ext(litExt, b, w, stream);
            check(cudaGetLastError(), "pack launch");
            totalLaunches += static_cast<uint64_t>(launches);
            packLaunches += static_cast<uint64_t>(launches);
            check(cudaMemcpyAsync(hFlags, w.flags, 64 * 4, cudaMemcpyDeviceToHost, stream), "D2H(pack flags)");
            check(cudaMemcpyAsync(hTot, w.blockTot + blocks * PACK_STREAMS, PACK_STREAMS * 4, cudaMemcpyDeviceToHost, stream), "D2H(pack totals)");
            if (nC) check(cudaMemcpyAsync(hSumLen, w.sumLen, nC * 8ull, cudaMemcpyDeviceToHost, stream), "D2H(pack sums)");
            check(cudaMemcpyAsync(hKOff, w.kOff, (nC + 1ull) * 4, cudaMemcpyDeviceToHost, stream), "D2H(kept offsets)");
            check(cudaMemcpyAsync(hStreamOff, w.streamOff, static_cast<uint64_t>(PACK_STREAMS) * (nC + 1) * 4, cudaMemcpyDeviceToHost, stream), "D2H(stream offsets)");
            if (wantLit) check(cudaMemcpyAsync(hLitCount, w.litFlag + nGaps, 4, cudaMemcpyDeviceToHost, stream), "D2H(literal range count)");
            check(cudaEventRecord(pkEv[2], stream), "event");
            check(cudaStreamSynchronize(stream), "pack sync");
            // a sweep that still changed something means the fixed point was not reached: run again with more sweeps
            // (slot 15 / 31 belongs to the last sweep alone: earlier sweeps share the other slots)
            const bool remOpen = nM && p.enableExt && hFlags[15];
            const bool claimOpen = nM && p.gapDepthOffset > 0 && hFlags[16 + 15];
            if (!remOpen && !claimOpen) break;
            // a dependency chain is at most as long as the list: more sweeps than matches cannot be needed
            if (attempt > 24 || static_cast<uint64_t>(std::max(remSweeps, claimSweeps)) > 2ull * nM + 4)
                throw MatcherError(MBGCG_E_CUDA, "stream packing did not reach its fixed point");
            if (remOpen) remSweeps *= 2;
            if (claimOpen) claimSweeps *= 2;
        }
        packSweeps[0] = std::max(packSweeps[0], remSweeps);
        packSweeps[1] = std::max(packSweeps[1], claimSweeps);

        // ---- streams -> host ----
        for (int k = 0; k < PACK_STREAMS; k++) {
            const uint64_t n = hTot[k];
            if (n > cap[k]) throw MatcherError(MBGCG_E_CUDA, "stream packing overran its output bound");
            if (outUsed[k] + n > hPackOutCap[k]) {
                const uint64_t want = align_up((outUsed[k] + n) + (outUsed[k] + n) / 4 + 4096, 1 << 16);
                uint8_t* grown = nullptr;
                check(cudaHostAlloc(reinterpret_cast<void**>(&grown), want, cudaHostAllocDefault), "cudaHostAlloc(pack streams)");
                if (outUsed[k]) std::memcpy(grown, hPackOut[k], outUsed[k]);
                if (hPackOut[k]) cudaFreeHost(hPackOut[k]);
                hPackOut[k] = grown;
                hPackOutCap[k] = want;
            }
            if (n) check(cudaMemcpyAsync(hPackOut[k] + outUsed[k], w.out[k], n, cudaMemcpyDeviceToHost, stream), "D2H(pack stream)");
        }
        const bool morePasses = last < nT;
        std::vector<KeptRec> keptTmp;
        if (morePasses && hKOff[nC]) {
            keptTmp.resize(hKOff[nC]);
            check(cudaMemcpyAsync(keptTmp.data(), w.kept, hKOff[nC] * sizeof(KeptRec), cudaMemcpyDeviceToHost, stream), "D2H(kept)");
        }
        std::vector<LitRange> litTmp;
        if (wantLit && *hLitCount) {
            litTmp.resize(*hLitCount);
            check(cudaMemcpyAsync(litTmp.data(), w.litRanges, litTmp.size() * sizeof(LitRange), cudaMemcpyDeviceToHost, stream), "D2H(literal ranges)");
        }
        check(cudaEventRecord(pkEv[3], stream), "event");
        check(cudaStreamSynchronize(stream), "pack D2H sync");
        for (LitRange r : litTmp) {
            r.text += static_cast<uint32_t>(first);
            packLitRanges.push_back(r);
        }
        for (int c = 0; c < nC; c++) {
            for (int k = 0; k < PACK_STREAMS; k++) packOff[k][first + c] = outUsed[k] + hStreamOff[static_cast<uint64_t>(k) * (nC + 1) + c];
            packKeptCounts[first + c] = hKOff[c + 1] - hKOff[c];
            packUnmatchedChars[first + c] = texts[first + c].len - hSumLen[c];  // :164-166,199 telescoped
        }
        for (int k = 0; k < PACK_STREAMS; k++) outUsed[k] += hTot[k];
        if (morePasses) {
            for (const KeptRec& r : keptTmp) packKeptHost.push_back(TextMatch{r.src, r.len, r.dst, 0});
        } else {
            packKeptLast = KeptSlice{w.kept, hKOff[nC]};
        }
        float ms = 0;
        check(cudaEventElapsedTime(&ms, pkEv[0], pkEv[1]), "event time");
        packMs[0] += ms;
        check(cudaEventElapsedTime(&ms, pkEv[1], pkEv[2]), "event time");
        packMs[1] += ms;
        check(cudaEventElapsedTime(&ms, pkEv[2], pkEv[3]), "event time");
        packMs[2] += ms;
        check(cudaEventElapsedTime(&ms, pkEv[0], pkEv[3]), "event time");
        packMs[3] += ms;
        matchBase += nM;
        first = last;
    }
    for (int k = 0; k < PACK_STREAMS; k++) packOff[k][nT] = outUsed[k];
}

void B200SparseEMMatcher::packKeptMatches(std::vector<TextMatch>& out) {
    check(cudaSetDevice(device), "cudaSetDevice");
    out = packKeptHost;
    if (packKeptLast.n) {
        std::vector<KeptRec> tmp(packKeptLast.n);
        check(cudaMemcpyAsync(tmp.data(), packKeptLast.dev, tmp.size() * sizeof(KeptRec), cudaMemcpyDeviceToHost, stream), "D2H(kept)");
        check(cudaStreamSynchronize(stream), "kept sync");
        for (const KeptRec& r : tmp) out.push_back(TextMatch{r.src, r.len, r.dst, 0});
    }
}

}  // namespace mbgc_b200
