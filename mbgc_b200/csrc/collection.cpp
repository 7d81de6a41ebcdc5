// mbgc_b200/csrc/collection.cpp -- see collection.hpp.
#include "collection.hpp"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cerrno>
#include <cstring>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/mbgc_b200.h"
#include "hostpack.hpp"
#include "kernels.h"

namespace mbgc_b200 {

namespace {
constexpr uint64_t PAD = 64, TEXT_GAP = 64, TEXT_TAIL = 8192;  // = matcher.cpp
constexpr uint64_t BOUNCE_BYTES = 8ull << 20;
constexpr int NSLOT = 3;  // ingest slots per replica: one in the window, one being ingested, one receiving images
inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }
inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
bool is_pinned(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}
}  // namespace

uint64_t packed_image_bound(uint64_t n) {
    const uint64_t maxRuns = n / 256 + 256;
    return sizeof(PackedImageHeader) + align_up((n + 3) / 4 + 16, 16) + maxRuns * sizeof(ExcRun) + 64;
}

uint64_t pack_image_into(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t cap) {
    if (n == 0 || cap < packed_image_bound(n) || (reinterpret_cast<uintptr_t>(dst) & 15)) return 0;
    const uint64_t maxRuns = n / 256 + 256;
    static thread_local std::vector<ExcRun> runs;
    runs.clear();
    PackedImageHeader h{};
    h.magic = PACKED_IMAGE_MAGIC;
    h.n = n;
    h.packedOff = sizeof(PackedImageHeader);
    h.packedBytes = align_up((n + 3) / 4 + 16, 16);
    if (!pack_fasta_image(src, n, dst + h.packedOff, runs, maxRuns)) return 0;
    h.runsOff = h.packedOff + h.packedBytes;
    h.nRuns = static_cast<uint32_t>(runs.size());
    h.firstByte = src[0];
    h.total = h.runsOff + runs.size() * sizeof(ExcRun);
    if (!runs.empty()) std::memcpy(dst + h.runsOff, runs.data(), runs.size() * sizeof(ExcRun));
    std::memcpy(dst, &h, sizeof(h));
    return h.total;
}

// The same straight from a file: read() into a buffer that stays in the core's cache, packed from there -- the page cache is
// read once, a quarter of it is written.  (An mmap of the file would save the copy into the buffer and let the packing
// loop's prefetch reach the page cache, but reader threads mapping and unmapping files serialise on the process's address
// space: measured twice -- page faults in round 2a, MAP_POPULATE in r02c: the 1000-file pass takes 0.19 s instead of 0.074 s.)
uint64_t pack_image_from_file(const char* path, uint8_t* dst, uint64_t cap, uint64_t* fileSize) {
    *fileSize = 0;
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) throw MatcherError(MBGCG_E_PARAM, std::string("cannot open ") + path);
    struct FdGuard {
        int fd;
        ~FdGuard() { ::close(fd); }
    } guard{fd};
    struct stat sb;
    if (::fstat(fd, &sb) != 0) throw MatcherError(MBGCG_E_PARAM, std::string("cannot stat ") + path);
    const uint64_t n = static_cast<uint64_t>(sb.st_size);
    *fileSize = n;
    if (n == 0 || cap < packed_image_bound(n) || (reinterpret_cast<uintptr_t>(dst) & 15)) return 0;
    const uint64_t maxRuns = n / 256 + 256;
    static thread_local std::vector<ExcRun> runs;
    static thread_local std::vector<uint8_t> buf;
    static const uint64_t PIECE = [] {  // a multiple of 64 that stays in the core's L2
        const char* e = std::getenv("MBGC_B200_READ_PIECE_KB");
        const uint64_t kb = e ? std::strtoull(e, nullptr, 10) : 256;
        return std::max<uint64_t>(4, std::min<uint64_t>(kb, 65536)) << 10;
    }();
    if (buf.size() < PIECE + 64) buf.resize(PIECE + 64);
    uint8_t* b = buf.data() + ((64 - (reinterpret_cast<uintptr_t>(buf.data()) & 63)) & 63);
    runs.clear();
    PackedImageHeader h{};
    h.magic = PACKED_IMAGE_MAGIC;
    h.n = n;
    h.packedOff = sizeof(PackedImageHeader);
    h.packedBytes = align_up((n + 3) / 4 + 16, 16);
    ImagePacker packer(dst + h.packedOff, runs, maxRuns);
    uint64_t done = 0;
    uint8_t first = 0;
    while (done < n) {
        const uint64_t want = std::min<uint64_t>(PIECE, n - done);
        uint64_t got = 0;
        while (got < want) {
            const ssize_t k = ::read(fd, b + got, want - got);
            if (k < 0 && errno == EINTR) continue;
            if (k <= 0) throw MatcherError(MBGCG_E_PARAM, std::string("short read on ") + path);
            got += static_cast<uint64_t>(k);
        }
        if (done == 0) first = b[0];
        if (!packer.add(b, want)) return 0;
        done += want;
    }
    if (!packer.finish()) return 0;
    h.runsOff = h.packedOff + h.packedBytes;
    h.nRuns = static_cast<uint32_t>(runs.size());
    h.firstByte = first;
    h.total = h.runsOff + runs.size() * sizeof(ExcRun);
    if (!runs.empty()) std::memcpy(dst + h.runsOff, runs.data(), runs.size() * sizeof(ExcRun));
    std::memcpy(dst, &h, sizeof(h));
    return h.total;
}

// ---------------------------------------------------------------------------------------------------------------------
// FastaBatchIngest
// ---------------------------------------------------------------------------------------------------------------------

void FastaBatchIngest::check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw MatcherError(MBGCG_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

void FastaBatchIngest::grow(uint8_t** p, uint64_t* cap, uint64_t need, const char* what, bool keep, uint64_t keepBytes) {
    if (need <= *cap) return;
    const uint64_t c = align_up(need + need / 8, 1 << 20);
    uint8_t* grown = nullptr;
    check(cudaMalloc(&grown, c), what);
    if (keep && *p && keepBytes) {
        check(cudaMemcpyAsync(grown, *p, keepBytes, cudaMemcpyDeviceToDevice, stream), what);
        check(cudaStreamSynchronize(stream), what);
    }
    cudaFree(*p);
    *p = grown;
    *cap = c;
}

FastaBatchIngest::FastaBatchIngest(int dev) : device(dev) {
    check(cudaSetDevice(device), "cudaSetDevice");
    check(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
    for (auto& e : ev) check(cudaEventCreate(&e), "cudaEventCreate");
    for (auto& e : bounceEv) check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
}

FastaBatchIngest::~FastaBatchIngest() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (auto& f : inflight)
        if (f.ev) cudaEventDestroy(f.ev);
    for (auto e : evPool) cudaEventDestroy(e);
    cudaFree(dIn);
    if (dPk) cudaFree(dPk);
    cudaFree(dWork);
    cudaFree(dText);
    cudaFree(dHdr);
    if (hWork) cudaFreeHost(hWork);
    if (hHdr) cudaFreeHost(hHdr);
    for (auto& b : hBounce)
        if (b) cudaFreeHost(b);
    for (auto& e : ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : bounceEv)
        if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
}

// The per-file loop of ingest.cpp (B200SparseEMMatcher::ingestFasta) for a batch of images in one device buffer: image f
// sits at a 16-byte aligned offset with 16 '\n' in front of it; a record ends at its file's end (k_fa_records clamps).
void FastaBatchIngest::begin(uint64_t expectedBytes) {
    check(cudaSetDevice(device), "cudaSetDevice");
    files.clear();
    contigs.clear();
    fileBytes.clear();
    hdrAt.clear();
    live.clear();
    spans.clear();
    cursor = 0;
    launches = 0;
    imageBytes = textBytes = pageableBytes = wireBytes = 0;
    ms[0] = ms[1] = 0;
    bounceUse = 0;
    if (!inflight.empty()) throw MatcherError(MBGCG_E_PARAM, "FASTA ingest: images of the previous batch were never reaped");
    grow(&dIn, &dInCap, 16 + expectedBytes + expectedBytes / 8 + 4096, "cudaMalloc(fasta images)");
    check(cudaEventRecord(ev[0], stream), "event");
}

void FastaBatchIngest::add(const HostImage& im) {
    check(cudaSetDevice(device), "cudaSetDevice");
    const int f = static_cast<int>(files.size());
    files.push_back(IngestedFile{});
    fileBytes.push_back(im.n);
    // only images that can hold a record travel (the others are decided here: kseq.h:239-240)
    if (im.n == 0) {                               // end of file before the first header: status -1, no records
        inflight.push_back(InFlight{nullptr, im.bytes});
        return;
    }
    const PackedImageHeader* ph = im.packed ? reinterpret_cast<const PackedImageHeader*>(im.bytes) : nullptr;
    if (ph && (ph->magic != PACKED_IMAGE_MAGIC || ph->n != im.n || ph->total < sizeof(PackedImageHeader)))
        throw MatcherError(MBGCG_E_PARAM, "a packed file image with a damaged header");
    if ((ph ? static_cast<char>(ph->firstByte) : im.bytes[0]) != '>') {  // not FASTA
        files[f].status = -3;
        inflight.push_back(InFlight{nullptr, im.bytes});
        return;
    }
    if (im.n >= (1ull << 40)) throw MatcherError(MBGCG_E_UNSUPPORTED, "FASTA image of 1 TiB or more");
    const uint64_t end = cursor + im.n;
    grow(&dIn, &dInCap, 16 + align_up(end, 16) + 16 + 64, "cudaMalloc(fasta images)", true, 16 + cursor);
    uint8_t* dst = dIn + 16 + cursor;
    if (ph) {
        // 4 bases per byte + exception runs over PCIe, the exact bytes of the file rebuilt here (kernels.cu: k_unpack_image,
        // k_patch_exceptions).  The slot's stream orders copy and kernels; the sync below hands the host memory back.
        const uint64_t body = ph->total - sizeof(PackedImageHeader);
        grow(&dPk, &dPkCap, body + 64, "cudaMalloc(packed image)");
        check(cudaMemcpyAsync(dPk, im.bytes + sizeof(PackedImageHeader), body, cudaMemcpyHostToDevice, stream), "H2D(packed image)");
        launches += launch_unpack_image(dPk + (ph->packedOff - sizeof(PackedImageHeader)), im.n,
                                        reinterpret_cast<const ExcRunDev*>(dPk + (ph->runsOff - sizeof(PackedImageHeader))), ph->nRuns, dst,
                                        stream);
        if (!is_pinned(im.bytes)) pageableBytes += ph->total;
        wireBytes += ph->total;
    } else if (is_pinned(im.bytes)) {
        wireBytes += im.n;
        check(cudaMemcpyAsync(dst, im.bytes, im.n, cudaMemcpyHostToDevice, stream), "H2D(fasta image)");
    } else {  // pageable: through the pinned bounce buffers, one being filled while the other is on the wire
        for (uint64_t o = 0; o < im.n; o += BOUNCE_BYTES) {
            const uint64_t m = std::min<uint64_t>(BOUNCE_BYTES, im.n - o);
            const int b = bounceUse & 1;
            if (!hBounce[b]) check(cudaHostAlloc(reinterpret_cast<void**>(&hBounce[b]), BOUNCE_BYTES, cudaHostAllocDefault), "cudaHostAlloc(bounce)");
            else if (bounceUse >= 2) check(cudaEventSynchronize(bounceEv[b]), "bounce sync");
            std::memcpy(hBounce[b], im.bytes + o, m);
            check(cudaMemcpyAsync(dst + o, hBounce[b], m, cudaMemcpyHostToDevice, stream), "H2D(fasta image)");
            check(cudaEventRecord(bounceEv[b], stream), "event");
            bounceUse++;
        }
        pageableBytes += im.n;
        wireBytes += im.n;
    }
    live.push_back(f);
    spans.push_back(cursor);
    spans.push_back(end);
    cursor = align_up(end, 16) + 16;
    imageBytes += im.n;
    // the caller gets its memory back when the copy is over (reap); a pageable image has been copied into the bounce buffers
    // by now and is free at once
    cudaEvent_t e = nullptr;
    if (ph || is_pinned(im.bytes)) {
        if (evPool.empty()) {
            check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
        } else {
            e = evPool.back();
            evPool.pop_back();
        }
        check(cudaEventRecord(e, stream), "event");
    }
    inflight.push_back(InFlight{e, im.bytes});
}

void FastaBatchIngest::reap(std::vector<const char*>& released, bool wait) {
    while (!inflight.empty()) {
        InFlight& f = inflight.front();
        if (f.ev) {
            if (wait) {
                check(cudaEventSynchronize(f.ev), "image copy sync");
            } else {
                const cudaError_t q = cudaEventQuery(f.ev);
                if (q == cudaErrorNotReady) break;
                check(q, "image copy");
            }
            evPool.push_back(f.ev);
        }
        released.push_back(f.bytes);
        inflight.pop_front();
    }
}

void FastaBatchIngest::finish() {
    check(cudaSetDevice(device), "cudaSetDevice");
    const int nFiles = static_cast<int>(files.size());
    const uint32_t nLive = static_cast<uint32_t>(live.size());
    if (nLive == 0) {
        grow(&dText, &dTextCap, PAD + TEXT_TAIL, "cudaMalloc(ingested texts)");
        return;
    }
    const uint64_t n = spans.back();                   // end of the last image
    const uint64_t inBytes = 16 + align_up(n, 16) + 64;
    grow(&dIn, &dInCap, inBytes, "cudaMalloc(fasta images)", true, 16 + n);
    uint8_t* in = dIn + 16;

    // descriptors: spans, then file ends
    const uint32_t blocks = static_cast<uint32_t>((n + FA_BLOCK_BYTES - 1) / FA_BLOCK_BYTES);
    const uint64_t oSpans = 0;
    const uint64_t oEnds = align_up(oSpans + nLive * 16ull, 256);
    const uint64_t oCnt = align_up(oEnds + nLive * 8ull, 256);
    const uint64_t cntBytes = align_up((blocks + 2ull) * 4, 256);
    const uint64_t fixedBytes = oCnt + cntBytes;
    grow(&dWork, &dWorkCap, fixedBytes + (1 << 20), "cudaMalloc(fasta work)");
    auto growHost = [&](uint8_t** p, uint64_t* cap, uint64_t need) {
        if (need <= *cap) return;
        if (*p) cudaFreeHost(*p);
        *p = nullptr;
        *cap = 0;
        const uint64_t c = align_up(need + need / 4, 1 << 16);
        check(cudaHostAlloc(reinterpret_cast<void**>(p), c, cudaHostAllocDefault), "cudaHostAlloc(fasta work)");
        *cap = c;
    };
    growHost(&hWork, &hWorkCap, oCnt + 64);
    std::memcpy(hWork + oSpans, spans.data(), nLive * 16ull);
    uint64_t* hEnds = reinterpret_cast<uint64_t*>(hWork + oEnds);
    for (uint32_t k = 0; k < nLive; k++) hEnds[k] = spans[2 * k + 1];
    check(cudaEventRecord(ev[1], stream), "event");
    check(cudaMemcpyAsync(dWork, hWork, oEnds + nLive * 8ull, cudaMemcpyHostToDevice, stream), "H2D(fasta descriptors)");
    const uint64_t* dSpans = reinterpret_cast<const uint64_t*>(dWork + oSpans);
    const uint64_t* dEnds = reinterpret_cast<const uint64_t*>(dWork + oEnds);
    int nl = launch_fa_filler(in, dSpans, nLive, inBytes - 16, stream);

    // ---- record starts ----
    uint32_t* dCnt = reinterpret_cast<uint32_t*>(dWork + oCnt);
    nl += launch_fa_count(in, n, dCnt, stream);
    uint32_t* hN = reinterpret_cast<uint32_t*>(hWork + oCnt);
    check(cudaMemcpyAsync(hN, dCnt + blocks, 4, cudaMemcpyDeviceToHost, stream), "D2H(record count)");
    check(cudaStreamSynchronize(stream), "ingest sync");
    const uint32_t nRecs = *hN;
    if (nRecs < nLive) throw MatcherError(MBGCG_E_CUDA, "FASTA ingest found fewer record starts than files");

    const uint64_t oStarts = fixedBytes;
    const uint64_t oRecs = align_up(oStarts + nRecs * 8ull, 256);
    const uint64_t oGather = align_up(oRecs + nRecs * sizeof(FaRecord), 256);
    const uint64_t oHdr = align_up(oGather + nRecs * sizeof(FaGather), 256);
    const uint64_t workBytes = oHdr + nRecs * 24ull + 256;
    grow(&dWork, &dWorkCap, workBytes, "cudaMalloc(fasta work)", true, fixedBytes);
    dCnt = reinterpret_cast<uint32_t*>(dWork + oCnt);
    dSpans = reinterpret_cast<const uint64_t*>(dWork + oSpans);
    dEnds = reinterpret_cast<const uint64_t*>(dWork + oEnds);
    uint64_t* dStarts = reinterpret_cast<uint64_t*>(dWork + oStarts);
    FaRecord* dRecs = reinterpret_cast<FaRecord*>(dWork + oRecs);
    FaGather* dGather = reinterpret_cast<FaGather*>(dWork + oGather);
    uint64_t* dHdrDesc = reinterpret_cast<uint64_t*>(dWork + oHdr);
    nl += launch_fa_starts(in, n, dCnt, dStarts, stream);
    nl += launch_fa_records(in, n, dStarts, nRecs, dRecs, stream, dEnds, nLive);
    growHost(&hWork, &hWorkCap, nRecs * (sizeof(FaRecord) + sizeof(FaGather) + 24) + 64);
    FaRecord* hRecs = reinterpret_cast<FaRecord*>(hWork);
    FaGather* hGather = reinterpret_cast<FaGather*>(hRecs + nRecs);
    uint64_t* hHdrDesc = reinterpret_cast<uint64_t*>(hGather + nRecs);
    check(cudaMemcpyAsync(hRecs, dRecs, nRecs * sizeof(FaRecord), cudaMemcpyDeviceToHost, stream), "D2H(records)");
    check(cudaStreamSynchronize(stream), "ingest sync");

    // ---- the reader's file-wide state, file by file, record by record (kseq.h:253-256, 260-265) ----
    uint64_t outCursor = PAD, chunks = 0, hdrBytes = 0;
    uint32_t accepted = 0, r = 0;
    for (uint32_t k = 0; k < nLive; k++) {
        const int f = live[k];
        IngestedFile& fi = files[f];
        const uint64_t fStart = spans[2 * k], fEnd = spans[2 * k + 1];
        fi.firstContig = static_cast<uint32_t>(contigs.size());
        uint64_t dna = 0, maxLast = 0;
        bool stopped = false;
        while (r < nRecs && hRecs[r].start < fEnd) {
            const FaRecord& rec = hRecs[r++];
            if (stopped) continue;
            if (rec.start + 1 >= fEnd) {  // '>' is the file's last byte: kseq.h:244 returns end of file, the record does not exist
                stopped = true;
                continue;
            }
            const uint64_t seqStart = rec.hdrEnd + 1;
            const uint64_t len = rec.regionEnd > seqStart ? rec.regionEnd - seqStart : 0;
            const uint64_t body = len - (rec.endsWithNl ? 1 : 0);
            const uint64_t line = rec.firstLine;
            bool bad = rec.bad != 0;
            uint64_t seqLen = 0, lastLen = 0;
            if (!bad && body) {
                const uint64_t fullLines = body / (line + 1);
                seqLen = body - fullLines;
                lastLen = body - fullLines * (line + 1);
                if (fullLines) {
                    if (dna == 0) dna = line;
                    else if (dna != line) bad = true;
                }
            }
            if (lastLen > maxLast) maxLast = lastLen;
            if (!bad && dna && dna < maxLast) bad = true;
            if (bad) {
                dna = UINT64_MAX;
                fi.status = -4;
                stopped = true;
                continue;
            }
            if (seqLen >= (1ull << 32)) throw MatcherError(MBGCG_E_UNSUPPORTED, "sequence of 4 GiB or more");
            hGather[accepted] = FaGather{seqStart, outCursor, seqLen, seqLen ? line : 0, chunks};
            const uint64_t hdrLen = std::min(rec.hdrEnd, fEnd) - (rec.start + 1);
            hHdrDesc[3 * accepted] = rec.start + 1;
            hHdrDesc[3 * accepted + 1] = hdrLen;
            hHdrDesc[3 * accepted + 2] = hdrBytes;
            hdrAt.push_back(hdrBytes);
            hdrBytes += hdrLen;
            contigs.push_back(IngestedContig{static_cast<uint32_t>(f), rec.start + 1 - fStart, hdrLen, outCursor, seqLen});
            const uint64_t next = align_up(outCursor + seqLen + TEXT_GAP, 16);
            chunks += (next - outCursor) / 16;
            outCursor = next;
            accepted++;
        }
        fi.dnaLineLen = dna;
        fi.maxLastLineLen = maxLast;
        fi.nContigs = static_cast<uint32_t>(contigs.size()) - fi.firstContig;
    }
    for (int f = 0; f < nFiles; f++)   // files that did not travel sit between their neighbours
        if (files[f].nContigs == 0) {
            uint32_t at = 0;
            for (int g = f - 1; g >= 0; g--)
                if (files[g].nContigs) { at = files[g].firstContig + files[g].nContigs; break; }
            files[f].firstContig = at;
        }

    // ---- sequences -> texts, header lines -> host ----
    textBytes = outCursor + TEXT_TAIL;
    grow(&dText, &dTextCap, textBytes, "cudaMalloc(ingested texts)");
    check(cudaMemsetAsync(dText, 0, PAD, stream), "memset(ingested texts, head)");
    check(cudaMemsetAsync(dText + outCursor, 0, TEXT_TAIL, stream), "memset(ingested texts, tail)");
    if (accepted) {
        check(cudaMemcpyAsync(dGather, hGather, accepted * sizeof(FaGather), cudaMemcpyHostToDevice, stream), "H2D(gather)");
        check(cudaMemcpyAsync(dHdrDesc, hHdrDesc, accepted * 24ull, cudaMemcpyHostToDevice, stream), "H2D(header descriptors)");
        nl += launch_fa_gather(in, dGather, accepted, chunks, dText, stream);
        grow(&dHdr, &dHdrCap, hdrBytes + 64, "cudaMalloc(headers)");
        growHost(&hHdr, &hHdrCap, hdrBytes + 64);
        nl += launch_fa_headers(in, dHdrDesc, accepted, dHdr, stream);
        if (hdrBytes) check(cudaMemcpyAsync(hHdr, dHdr, hdrBytes, cudaMemcpyDeviceToHost, stream), "D2H(headers)");
    } else
        growHost(&hHdr, &hHdrCap, 64);
    check(cudaGetLastError(), "ingest launch");
    check(cudaEventRecord(ev[2], stream), "event");
    check(cudaStreamSynchronize(stream), "ingest sync");
    check(cudaEventElapsedTime(&ms[0], ev[1], ev[2]), "event time");
    check(cudaEventElapsedTime(&ms[1], ev[0], ev[2]), "event time");
    launches = static_cast<uint64_t>(nl);
}

// ---------------------------------------------------------------------------------------------------------------------
// PinnedArena
// ---------------------------------------------------------------------------------------------------------------------

PinnedArena::~PinnedArena() {
    for (auto& s : slabs)
        if (s.base) cudaFreeHost(s.base);
}

void* PinnedArena::alloc(uint64_t n) {
    n = align_up(std::max<uint64_t>(n, 1), 256);
    if (n > slabBytes) return nullptr;
    std::lock_guard<std::mutex> g(mu);
    for (auto& s : slabs)
        if (s.open && s.used + n <= slabBytes) {
            void* p = s.base + s.used;
            s.used += n;
            s.live++;
            return p;
        }
    for (auto& s : slabs)          // the open slab is full: close it; an empty closed slab starts over
        if (s.open) s.open = false;
    for (auto& s : slabs)
        if (!s.open && s.live == 0) {
            s.open = true;
            s.used = n;
            s.live = 1;
            return s.base;
        }
    if (static_cast<int>(slabs.size()) >= maxSlabs) return nullptr;
    uint8_t* base = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&base), slabBytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    slabs.push_back(Slab{base, n, 1, true});
    return base;
}

bool PinnedArena::owns(const void* p) {
    std::lock_guard<std::mutex> g(mu);
    const uint8_t* q = static_cast<const uint8_t*>(p);
    for (auto& s : slabs)
        if (q >= s.base && q < s.base + slabBytes) return true;
    return false;
}

bool PinnedArena::release(const void* p) {
    std::lock_guard<std::mutex> g(mu);
    const uint8_t* q = static_cast<const uint8_t*>(p);
    for (auto& s : slabs)
        if (q >= s.base && q < s.base + slabBytes) {
            if (s.live > 0) s.live--;
            if (s.live == 0 && !s.open) s.used = 0;
            return true;
        }
    return false;
}

// ---------------------------------------------------------------------------------------------------------------------
// ReplicaSet
// ---------------------------------------------------------------------------------------------------------------------

ReplicaSet::ReplicaSet(uint64_t refLengthLimit, uint32_t targetMatchLength, int k1, int k2, int skipMargin, const int* devices,
                       int ndev) {
    if (ndev < 1 || ndev > 64) throw MatcherError(MBGCG_E_PARAM, "between 1 and 64 devices");
    for (int i = 0; i < ndev; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) throw MatcherError(MBGCG_E_PARAM, "a device may hold one replica only");
    rep.resize(static_cast<size_t>(ndev), nullptr);
    workers.resize(static_cast<size_t>(ndev));
    try {
        // contexts and the gigabytes of reference / table come up on all devices at once
        std::vector<std::thread> th;
        std::vector<std::exception_ptr> errs(static_cast<size_t>(ndev));
        for (int r = 0; r < ndev; r++)
            th.emplace_back([&, r] {
                try {
                    rep[r] = new B200SparseEMMatcher(refLengthLimit, targetMatchLength, k1, k2, skipMargin, devices[r]);
                } catch (...) {
                    errs[r] = std::current_exception();
                }
            });
        for (auto& t : th) t.join();
        for (auto& e : errs)
            if (e) std::rethrow_exception(e);
        for (int r = 0; r < ndev; r++)   // peers read each other's resident texts (extension broadcast)
            for (int q = 0; q < ndev; q++) {
                if (q == r) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, devices[r], devices[q]);
                if (can) {
                    cudaSetDevice(devices[r]);
                    if (cudaDeviceEnablePeerAccess(devices[q], 0) != cudaSuccess) cudaGetLastError();
                }
            }
        for (int r = 0; r < ndev; r++) {
            workers[r].reset(new Worker());
            workers[r]->th = std::thread([this, r] { workerLoop(r); });
        }
    } catch (...) {
        for (auto& w : workers)
            if (w && w->th.joinable()) {
                { std::lock_guard<std::mutex> g(w->mu); w->quit = true; }
                w->cv.notify_all();
                w->th.join();
            }
        for (auto* m : rep) delete m;
        throw;
    }
}

ReplicaSet::~ReplicaSet() {
    for (auto& w : workers)
        if (w && w->th.joinable()) {
            { std::lock_guard<std::mutex> g(w->mu); w->quit = true; }
            w->cv.notify_all();
            w->th.join();
        }
    slotStore.clear();
    pinned.reset();
    for (auto* m : rep) delete m;
}

void ReplicaSet::workerLoop(int r) {
    Worker& w = *workers[r];
    cudaSetDevice(rep[r]->deviceIndex());
    std::unique_lock<std::mutex> lk(w.mu);
    for (;;) {
        w.cv.wait(lk, [&] { return w.quit || w.job; });
        if (w.quit) return;
        std::function<void()> job = std::move(w.job);
        w.job = nullptr;
        lk.unlock();
        std::exception_ptr err;
        try {
            job();
        } catch (...) {
            err = std::current_exception();
        }
        lk.lock();
        w.err = err;
        w.busy = false;
        w.cv.notify_all();
    }
}

void ReplicaSet::parallel(uint64_t which, const std::function<void(int)>& f) {
    const int n = size();
    int only = -1, cnt = 0;
    for (int r = 0; r < n; r++)
        if (which >> r & 1) { only = r; cnt++; }
    if (cnt == 0) return;
    if (cnt == 1) {  // nothing to overlap: run on the calling thread
        f(only);
        return;
    }
    for (int r = 0; r < n; r++) {
        if (!(which >> r & 1)) continue;
        Worker& w = *workers[r];
        std::lock_guard<std::mutex> g(w.mu);
        w.busy = true;
        w.err = nullptr;
        w.job = [&f, r] { f(r); };
        w.cv.notify_all();
    }
    std::exception_ptr first;
    for (int r = 0; r < n; r++) {
        if (!(which >> r & 1)) continue;
        Worker& w = *workers[r];
        std::unique_lock<std::mutex> lk(w.mu);
        w.cv.wait(lk, [&] { return !w.busy; });
        if (w.err && !first) first = w.err;
    }
    if (first) std::rethrow_exception(first);
}

void ReplicaSet::loadRef(const char* text, uint64_t len, bool rc, bool addSep, char sep) {
    parallel((size() >= 64 ? ~0ull : (1ull << size()) - 1), [&](int r) { rep[r]->loadRef(text, len, rc, addSep, sep); });
}

void ReplicaSet::loadRefFromReplica(int owner, const void* dText, uint64_t len, bool rc, bool addSep, char sep) {
    parallel((size() >= 64 ? ~0ull : (1ull << size()) - 1), [&](int r) {
        B200SparseEMMatcher* m = rep[r];
        if (m->hostMirror() && len) throw MatcherError(MBGCG_E_PARAM, "extensions from device memory need the host mirror off");
        if (r == owner || len == 0) {
            m->loadRefDevice(dText, nullptr, len, rc, addSep, sep);
            return;
        }
        void* stage = m->refStaging(len);
        cudaError_t e = cudaMemcpyPeerAsync(stage, m->deviceIndex(), dText, rep[owner]->deviceIndex(), len,
                                            static_cast<cudaStream_t>(m->streamHandle()));
        if (e != cudaSuccess) throw MatcherError(MBGCG_E_CUDA, std::string("peer copy of a reference extension: ") + cudaGetErrorString(e));
        m->loadRefDevice(stage, nullptr, len, rc, addSep, sep);
    });
}

void ReplicaSet::loadRefConcatFromReplica(int owner, const B200SparseEMMatcher::StagePiece* pieces, size_t n, bool rc,
                                          bool addSep, char sep) {
    uint64_t total = 0;
    void* src = rep[owner]->stageConcat(pieces, n, &total);
    parallel((size() >= 64 ? ~0ull : (1ull << size()) - 1), [&](int r) {
        B200SparseEMMatcher* m = rep[r];
        if (m->hostMirror() && total) throw MatcherError(MBGCG_E_PARAM, "extensions from device memory need the host mirror off");
        if (r == owner || total == 0) {
            m->loadRefDevice(src, nullptr, total, rc, addSep, sep);
            return;
        }
        void* stage = m->refStaging(total);
        cudaError_t e = cudaMemcpyPeerAsync(stage, m->deviceIndex(), src, rep[owner]->deviceIndex(), total,
                                            static_cast<cudaStream_t>(m->streamHandle()));
        if (e != cudaSuccess) throw MatcherError(MBGCG_E_CUDA, std::string("peer copy of a reference extension: ") + cudaGetErrorString(e));
        m->loadRefDevice(stage, nullptr, total, rc, addSep, sep);
    });
}

void ReplicaSet::loadSeparator(char sep) {
    for (auto* m : rep) m->loadSeparator(sep);
}

uint64_t ReplicaSet::acquireLock() {
    uint64_t v = 0;
    for (size_t r = 0; r < rep.size(); r++) {
        const uint64_t x = rep[r]->acquireWorkerMatchingLockPos();
        if (r == 0) v = x;
    }
    return v;
}

void ReplicaSet::releaseLock(uint64_t v) {
    for (auto* m : rep) m->releaseWorkerMatchingLockPos(v);
}

void ReplicaSet::setHostMirror(bool on) {
    for (auto* m : rep) m->setHostMirror(on);
}

void ReplicaSet::syncStateFromPrimary() {
    const B200SparseEMMatcher* p = rep[0];
    for (size_t r = 1; r < rep.size(); r++) {
        B200SparseEMMatcher* m = rep[r];
        if (m->pos1 != p->pos1 || m->reachedRefLengthCount != p->reachedRefLengthCount) m->invalidateMemo();
        m->pos1 = p->pos1;
        m->reachedRefLengthCount = p->reachedRefLengthCount;
        m->circularBuffer = p->circularBuffer;
        m->swSize = p->swSize;
        m->swEnd = p->swEnd;
        m->samplingPos = p->samplingPos;
    }
}

PinnedArena& ReplicaSet::arena() {
    std::lock_guard<std::mutex> g(arenaMu);
    if (!pinned) {
        uint64_t slab = 16ull << 20;
        int maxSlabs = 2 * size() + 6;
        if (const char* e = std::getenv("MBGC_B200_ARENA_SLAB_MB")) slab = std::max<uint64_t>(1, std::strtoull(e, nullptr, 10)) << 20;
        if (const char* e = std::getenv("MBGC_B200_ARENA_SLABS")) maxSlabs = std::max(0, std::atoi(e));
        pinned.reset(new PinnedArena(slab, maxSlabs));
    }
    return *pinned;
}

// ---------------------------------------------------------------------------------------------------------------------
// SequentialPass
// ---------------------------------------------------------------------------------------------------------------------

SequentialPass::SequentialPass(ReplicaSet& rs_, const SeqConfig& c, const uint64_t* loadedPosArr, uint64_t nLoaded)
    : rs(rs_), cfg(c), loaded(loadedPosArr, loadedPosArr + nLoaded), slots(rs_.ingestSlots()) {
    const int nd = rs.size();
    if (const char* e = std::getenv("MBGC_B200_CHUNK_MB")) cfg.chunkBytes = std::max<uint64_t>(1, std::strtoull(e, nullptr, 10)) << 20;
    if (const char* e = std::getenv("MBGC_B200_WINDOW")) cfg.maxWindowContigs = static_cast<uint32_t>(std::max(1, std::atoi(e)));
    rs.setHostMirror(false);  // extensions come from device memory; nothing on this path reads getRef()
    // MBGC_Encoder::processLiteral (row a12, mbgcg_set_literal_ref_ext on the handle): every replica lists the literal runs
    // of its contigs; a contig that is not loaded whole extends the reference with them (MGMP.cpp:281-286)
    const LitExtParams le = rs.primary()->literalRefExt();
    litExtOn = le.minLen != UINT64_MAX;
    if (litExtOn && cfg.parallelSchedule)
        throw MatcherError(MBGCG_E_UNSUPPORTED, "literal reference extensions belong to the sequential schedule (the reference's "
                                                "parallel loop reverse-complements the tail of targetRefExtensions, MGMP.cpp:392-397)");
    for (int r = 0; r < nd; r++)
        rs.at(r)->setLiteralRefExt(le.minLen, le.beforeAfter, static_cast<uint64_t>(std::max(0, cfg.unmatchedFractionFactor)));
    if (slots.size() < static_cast<size_t>(nd) * NSLOT) slots.resize(static_cast<size_t>(nd) * NSLOT);
    slotFree.assign(static_cast<size_t>(nd) * NSLOT, true);
    stageQ.resize(static_cast<size_t>(nd) * NSLOT);
    runState.resize(static_cast<size_t>(nd));
    curWindow = cfg.maxWindowContigs;
    for (int q = 0; q < nd * NSLOT; q++) stagers.emplace_back([this, q] { stagerLoop(q); });  // one per ingest slot
}

SequentialPass::~SequentialPass() {
    {
        std::lock_guard<std::mutex> g(mu);
        quit = true;
    }
    cv.notify_all();
    for (auto& t : stagers) t.join();
    // whatever was pushed but never consumed: arena images go back
    for (auto& im : pending) rs.arena().release(im.bytes);
}

void SequentialPass::pushFile(const char* bytes, uint64_t n, bool packed) {
    std::lock_guard<std::mutex> g(mu);
    if (inputDone) throw MatcherError(MBGCG_E_PARAM, "pushFile after finishInput");
    pending.push_back(HostImage{bytes, n, packed});
    pendingBytes += n;
    pushedFiles++;
    dispatch();
}

void SequentialPass::finishInput() {
    std::lock_guard<std::mutex> g(mu);
    inputDone = true;
    dispatch();
}

bool SequentialPass::wantsInput() {
    std::lock_guard<std::mutex> g(mu);
    if (inputDone) return false;
    // keep about two windows of images ahead of the commit cursor
    return pendingBytes + stagedBytes < 3 * cfg.chunkBytes * static_cast<uint64_t>(rs.size());
}

// (mu held) pending files -> the chunk being formed on the replica whose turn it is; every file starts its trip to the
// device at once, the chunk's kernels run when it is full
void SequentialPass::dispatch() {
    while (!pending.empty()) {
        if (!forming) {
            const int r = static_cast<int>(chunkSeq % static_cast<uint64_t>(rs.size()));
            int slot = -1;
            for (int s = 0; s < NSLOT; s++)
                if (slotFree[r * NSLOT + s]) { slot = s; break; }
            if (slot < 0) break;
            forming = std::make_shared<Chunk>();
            forming->replica = r;
            forming->slot = slot;
            forming->firstFile = dispatchedFiles;
            slotFree[r * NSLOT + slot] = false;
            chunkSeq++;
            staged.push_back(forming);
            stageQ[r * NSLOT + slot].push_back(StageMsg{StageMsg::BEGIN, forming, HostImage{nullptr, cfg.chunkBytes}});
        }
        const HostImage im = pending.front();
        pending.pop_front();
        pendingBytes -= im.n;
        stagedBytes += im.n;
        forming->nFiles++;
        forming->bytes += im.n;
        dispatchedFiles++;
        stageQ[forming->replica * NSLOT + forming->slot].push_back(StageMsg{StageMsg::ADD, forming, im});
        const uint64_t maxFiles = cfg.parallelSchedule
            ? static_cast<uint64_t>(std::max(1, cfg.allowedTargetsOutrun / rs.size())) : UINT64_MAX;
        if (forming->bytes >= cfg.chunkBytes || forming->nFiles >= maxFiles) {
            forming->closed = true;
            stageQ[forming->replica * NSLOT + forming->slot].push_back(StageMsg{StageMsg::FINISH, forming, HostImage{nullptr, 0}});
            forming.reset();
        }
    }
    if (inputDone && forming && pending.empty()) {
        forming->closed = true;
        stageQ[forming->replica * NSLOT + forming->slot].push_back(StageMsg{StageMsg::FINISH, forming, HostImage{nullptr, 0}});
        forming.reset();
    }
    cv.notify_all();
}

// one thread per ingest slot: while one slot's chunk runs its kernels and host round trips (finish), the replica's other
// slot already receives the images of the next chunk
void SequentialPass::stagerLoop(int q) {
    const int r = q / NSLOT;
    cudaSetDevice(rs.at(r)->deviceIndex());
    for (;;) {
        StageMsg msg;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return quit || !stageQ[q].empty(); });
            if (quit) {
                for (auto& m : stageQ[q])
                    if (m.type == StageMsg::ADD) rs.arena().release(m.image.bytes);
                return;
            }
            msg = stageQ[q].front();
            stageQ[q].pop_front();
        }
        Chunk& c = *msg.chunk;
        const double t0 = now_ms();
        bool added = false;
        std::vector<const char*> released;
        try {
            auto& slot = slots[r * NSLOT + c.slot];
            if (!slot) slot.reset(new FastaBatchIngest(rs.at(r)->deviceIndex()));
            if (msg.type == StageMsg::BEGIN) {
                slot->begin(msg.image.n);
            } else if (msg.type == StageMsg::ADD) {
                if (!c.err) {
                    slot->add(msg.image);
                    added = true;
                }
            } else if (!c.err) {
                slot->reap(released, true);
                slot->finish();
            }
            if (added) slot->reap(released, false);
        } catch (...) {
            if (!c.err) c.err = std::current_exception();
            try {  // whatever is still on its way: wait for it, the memory goes back either way
                auto& slot = slots[r * NSLOT + c.slot];
                if (slot) slot->reap(released, true);
            } catch (...) {
            }
        }
        if (msg.type == StageMsg::ADD && !added) released.push_back(msg.image.bytes);  // never reached the slot
        for (const char* p : released) rs.arena().release(p);  // arena images are ours to give back; others stay with the caller
        std::lock_guard<std::mutex> g(mu);
        st.ingestMs += now_ms() - t0;
        if (msg.type == StageMsg::FINISH) {
            if (!c.err) {
                auto& slot = slots[r * NSLOT + c.slot];
                st.imageBytes += slot->imageBytes;
                st.textBytes += slot->textBytes;
                st.ingestLaunches += slot->launches;
                st.pageableBytes += slot->pageableBytes;
                st.wireBytes += slot->wireBytes;
            }
            st.chunks++;
            c.ready = true;
            cv.notify_all();
        }
    }
}

// next window = the next (up to) one chunk per replica, in file order
bool SequentialPass::startWindow() {
    std::unique_lock<std::mutex> lk(mu);
    size_t take = 0;
    while (take < staged.size() && take < static_cast<size_t>(rs.size()) && staged[take]->closed) take++;
    if (take == 0) return false;
    // The batch-frozen schedule matches a window against the state at its start: WHICH targets share a window is part of the
    // result, so it must not depend on how fast the files arrive.  A window of fewer chunks than replicas starts only when no
    // further chunk can close before it: the input is over, or the pass does not want more of it (a function of the file
    // sizes alone).
    if (cfg.parallelSchedule && take < static_cast<size_t>(rs.size())) {
        const bool noMoreChunks = inputDone && pending.empty() && !forming;
        const bool wantsMore = !inputDone && pendingBytes + stagedBytes < 3 * cfg.chunkBytes * static_cast<uint64_t>(rs.size());
        if (!noMoreChunks && (wantsMore || !pending.empty())) return false;
    }
    win.clear();
    const double t0 = now_ms();
    for (size_t k = 0; k < take; k++) {
        std::shared_ptr<Chunk> c = staged.front();
        cv.wait(lk, [&] { return c->ready; });
        staged.pop_front();
        win.push_back(c);
    }
    st.waitStageMs += now_ms() - t0;
    lk.unlock();
    for (auto& c : win)
        if (c->err) std::rethrow_exception(c->err);
    items.clear();
    for (size_t w = 0; w < win.size(); w++) {
        FastaBatchIngest& ing = *slots[win[w]->replica * NSLOT + win[w]->slot];
        for (size_t f = 0; f < ing.files.size(); f++) {
            const IngestedFile& fi = ing.files[f];
            for (uint32_t k = 0; k < fi.nContigs; k++) items.push_back(Item{false, static_cast<int>(w), fi.firstContig + k});
            items.push_back(Item{true, static_cast<int>(w), static_cast<uint32_t>(f)});
        }
    }
    if (cfg.parallelSchedule) {  // every target takes its lock at the state it is matched against (MGMP.cpp:353-358)
        for (auto& c : win) {
            FastaBatchIngest& ing = *slots[c->replica * NSLOT + c->slot];
            c->locks.assign(ing.files.size(), 0);
            for (size_t f = 0; f < ing.files.size(); f++) c->locks[f] = rs.acquireLock();
        }
        windowFirstTarget = st.files;
    }
    cursor = 0;
    winActive = true;
    st.windows++;
    return true;
}

void SequentialPass::releaseWindow() {
    std::lock_guard<std::mutex> g(mu);
    for (auto& c : win) {
        slotFree[c->replica * NSLOT + c->slot] = true;
        stagedBytes -= c->bytes;
    }
    win.clear();
    items.clear();
    winActive = false;
    dispatch();
}

// match + pack the not yet valid contigs items[from ...) against the current state, at most curWindow of them and at
// most one pass of text per replica; every replica works on its own chunk
void SequentialPass::runFrom(size_t from) {
    uint64_t maxPassBytes = 1024ull << 20;
    if (const char* e = std::getenv("MBGC_B200_MAX_PASS_BYTES")) maxPassBytes = std::max<uint64_t>(1, std::strtoull(e, nullptr, 10));
    const int nd = rs.size();
    for (auto& rsn : runState) { rsn.texts.clear(); rsn.offs.clear(); rsn.locks.clear(); }
    std::vector<uint64_t> bytes(static_cast<size_t>(nd), 0);
    uint32_t taken = 0;
    uint64_t which = 0;
    for (size_t i = from; i < items.size() && taken < curWindow; i++) {
        Item& it = items[i];
        if (it.fileEnd) continue;
        const int r = win[it.chunk]->replica;
        FastaBatchIngest& ing = *slots[r * NSLOT + win[it.chunk]->slot];
        const IngestedContig& ic = ing.contigs[it.contig];
        if (!runState[r].texts.empty() && bytes[r] + ic.seqLen > maxPassBytes) break;
        bytes[r] += ic.seqLen;
        it.resIndex = static_cast<uint32_t>(runState[r].texts.size());
        it.valid = true;
        runState[r].texts.push_back(TextView{nullptr, ic.seqLen});
        runState[r].offs.push_back(ic.textOff);
        if (cfg.parallelSchedule) runState[r].locks.push_back(win[it.chunk]->locks[ic.file]);
        which |= 1ull << r;
        taken++;
    }
    std::vector<double> tm(static_cast<size_t>(nd) * 2, 0.0);
    std::vector<uint64_t> launches(static_cast<size_t>(nd), 0);
    rs.parallel(which, [&](int r) {
        B200SparseEMMatcher* m = rs.at(r);
        // the window's contigs of this replica all live in one ingest slot (one chunk per replica and window)
        int slot = -1;
        for (auto& c : win)
            if (c->replica == r) slot = c->slot;
        FastaBatchIngest& ing = *slots[r * NSLOT + slot];
        const uint64_t l0 = m->totalLaunches;
        const double t0 = now_ms();
        const uint64_t* locks = cfg.parallelSchedule ? runState[r].locks.data() : nullptr;
        m->matchBatch(runState[r].texts, locks, cfg.minMatchLength, true, ing.dTexts(), runState[r].offs.data(), true);
        const double t1 = now_ms();
        m->packBatch(cfg.pack, runState[r].texts, nullptr, nullptr, locks, loaded.data(), loaded.size());
        tm[r * 2] = t1 - t0;
        tm[r * 2 + 1] = now_ms() - t1;
        launches[r] = m->totalLaunches - l0;
    });
    double mMax = 0, pMax = 0;
    for (int r = 0; r < nd; r++) {
        mMax = std::max(mMax, tm[r * 2]);
        pMax = std::max(pMax, tm[r * 2 + 1]);
        st.launches += launches[r];
    }
    st.matchMs += mMax;
    st.packMs += pMax;
    st.runs++;
    committedThisRun = 0;
    if (cfg.parallelSchedule) windowFirstTarget = st.files;   // these results saw every target finalized so far
}

void SequentialPass::next(SeqEvent& ev) {
    for (;;) {
        if (!evQ.empty()) {
            ev = evQ.front();
            evQ.pop_front();
            return;
        }
        if (!winActive) {
            if (!startWindow()) {
                std::lock_guard<std::mutex> g(mu);
                ev = SeqEvent{};
                if (inputDone && pending.empty() && staged.empty() && !forming) {
                    ev.type = SeqEvent::DONE;
                    ev.loadedRefLength = rs.primary()->getLoadedRefLength();
                } else
                    ev.type = SeqEvent::NEED_INPUT;
                return;
            }
            continue;
        }
        if (cursor >= items.size()) {
            releaseWindow();
            continue;
        }
        Item& it = items[cursor];
        Chunk& ch = *win[it.chunk];
        FastaBatchIngest& ing = *slots[ch.replica * NSLOT + ch.slot];
        B200SparseEMMatcher* pm = rs.primary();
        if (it.fileEnd) {
            const IngestedFile& fi = ing.files[it.contig];
            SeqEvent e{};
            e.type = SeqEvent::FILE_END;
            e.window = st.windows;
            e.file = ch.firstFile + it.contig;
            e.status = fi.status;
            e.nContigs = fi.nContigs;
            e.dnaLineLen = fi.dnaLineLen == UINT64_MAX ? 0 : fi.dnaLineLen;  // KSEQ_DNA_LINE_LENGTH, MGMP.cpp:12-14
            if (cfg.parallelSchedule) {
                // finalizeParallelProcessingOfTarget (MGMP.cpp:433-468 with the encoder's hook, MBGC_Encoder.cpp:542-564):
                // the target's extension goes into the reference as ONE text, then separator, extension size, lock release
                const uint64_t lock = ch.locks[it.contig];
                if (fi.status >= -2) {
                    const uint64_t start = pm->getLoadedRefLength();
                    if (!fileExt.empty()) {
                        const double t0 = now_ms();
                        loadRefPieces(ch.replica, fileExt, !cfg.rcDisabled && !cfg.contigsIndividuallyReversed);
                        st.loadRefMs += now_ms() - t0;
                        st.extensions++;
                    }
                    if (cfg.separatorPerFile) {
                        rs.loadSeparator(cfg.separator);
                        e.refExtSize = pm->getLoadedRefLength() - start;
                        loaded.push_back(pm->getLoadedRefLength());
                    }
                }
                e.lockPos = lock;
                rs.releaseLock(lock);
                fileExt.clear();
            } else {
                // processTargetMeta + processAfterTargetWithParallelIO (MBGC_Encoder.cpp:493-509): separator, the size of
                // this file's reference extension, and the sliding-window position the decoder will replay
                if (!fileOpen) fileStartLoaded = pm->getLoadedRefLength();  // a file without contigs
                if (fi.status >= -2) {
                    if (cfg.separatorPerFile) {
                        rs.loadSeparator(cfg.separator);
                        e.refExtSize = pm->getLoadedRefLength() - fileStartLoaded;
                        loaded.push_back((loaded.empty() ? 0 : loaded.back()) + e.refExtSize);
                    }
                    e.lockPos = rs.acquireLock();
                    rs.releaseLock(e.lockPos);
                }
            }
            e.loadedRefLength = pm->getLoadedRefLength();
            fileOpen = false;
            st.files++;
            evQ.push_back(e);
            cursor++;
            continue;
        }
        if (!it.valid) {
            runFrom(cursor);
            continue;
        }
        const IngestedContig& ic = ing.contigs[it.contig];
        B200SparseEMMatcher* m = rs.at(ch.replica);
        const uint64_t unmatched = m->packUnmatched()[it.resIndex];
        if (cfg.parallelSchedule) {
            // the dissimilar-contig rule of MBGC_Encoder::processMatches (:202-205; MGMP_Params.h:193-196): a contig that
            // looks dissimilar against a state more than allowedTargetsOutrunForDissimilarContigs targets old is matched
            // again once the targets in front of it are finalized (MGMP.cpp:382-388) -- here: now, they all are
            const uint64_t target = ch.firstFile + ic.file;
            const int factor = cfg.unmatchedFractionFactor / std::max(1, cfg.unmatchedFractionFactorTweakForDissimilarContigs);
            if (windowFirstTarget + static_cast<uint64_t>(cfg.allowedTargetsOutrunForDissimilarContigs) < target &&
                ic.seqLen > cfg.minimalLengthForDissimilarContigs && unmatched * static_cast<uint64_t>(factor) > ic.seqLen) {
                for (size_t i = cursor; i < items.size(); i++)
                    if (!items[i].fileEnd && items[i].valid) {
                        items[i].valid = false;
                        st.rerunContigs++;
                    }
                runFrom(cursor);
                continue;
            }
        }
        // ---- commit one contig (MultipleGenomeMatchingProcessor.cpp:274-287 / :379-398) ----
        if (!fileOpen) {
            fileStartLoaded = pm->getLoadedRefLength();
            fileOpen = true;
        }
        SeqEvent e{};
        e.type = SeqEvent::CONTIG;
        e.window = st.windows;
        e.file = ch.firstFile + ic.file;
        e.hdr = ing.header(it.contig);
        e.hdrLen = ic.hdrLen;
        e.seqLen = ic.seqLen;
        e.nMatches = m->batchCounts()[it.resIndex];
        e.kept = m->packKept()[it.resIndex];
        e.unmatched = unmatched;
        e.contigLock = cfg.parallelSchedule ? ch.locks[ic.file] : UINT64_MAX;
        for (int s = 0; s < PACK_STREAMS; s++) {
            const auto& off = m->packStreamOff(s);
            e.seg[s] = m->packStream(s) + off[it.resIndex];
            e.segLen[s] = off[it.resIndex + 1] - off[it.resIndex];
        }
        const bool extend = e.unmatched * static_cast<uint64_t>(cfg.unmatchedFractionFactor) > ic.seqLen;      // MGMP_Params.h:179-187
        const bool extendRC = !cfg.rcDisabled && e.unmatched * static_cast<uint64_t>(cfg.unmatchedFractionRCFactor) > ic.seqLen;  // :189-191
        e.extended = extend;
        e.extendedRC = extend && extendRC;
        st.contigs++;
        committedThisRun++;
        cursor++;
        evQ.push_back(e);
        if (cfg.parallelSchedule) {
            // targetRefExtensions[i]: the contig, and right behind it its reverse complement (MGMP.cpp:389-398)
            if (extend) fileExt.push_back(ExtPiece{ing.dTexts() + ic.textOff, ic.seqLen, false});
            if (extend && extendRC && cfg.contigsIndividuallyReversed) fileExt.push_back(ExtPiece{ing.dTexts() + ic.textOff, ic.seqLen, true});
            continue;
        }
        // processLiteral: the long literal runs of a contig that is not loaded whole (targetRefExtensions[0], MGMP.cpp:281-282)
        std::vector<ExtPiece> litPieces;
        if (!extend && litExtOn) {
            const auto& lr = m->packLiteralRefExt();  // ordered by text, then by position
            auto lo = std::lower_bound(lr.begin(), lr.end(), it.resIndex, [](const LitRange& a, uint32_t t) { return a.text < t; });
            for (; lo != lr.end() && lo->text == it.resIndex; ++lo)
                litPieces.push_back(ExtPiece{ing.dTexts() + ic.textOff + lo->start, lo->len, false});
        }
        if (extend || !litPieces.empty()) {
            const double t0 = now_ms();
            if (extend)
                rs.loadRefFromReplica(ch.replica, ing.dTexts() + ic.textOff, ic.seqLen, extendRC, cfg.refRegionSeparators, cfg.separator);
            else
                loadRefPieces(ch.replica, litPieces, extendRC);
            st.loadRefMs += now_ms() - t0;
            st.extensions++;
            // everything matched against the old state and not yet committed runs again; speculation that fails is wasted
            // work, so the window follows the commit rate (as mbgc_b200/dist.py does)
            for (size_t i = cursor; i < items.size(); i++)
                if (!items[i].fileEnd && items[i].valid) {
                    items[i].valid = false;
                    st.rerunContigs++;
                }
            curWindow = std::max<uint32_t>(static_cast<uint32_t>(rs.size()), std::min<uint32_t>(cfg.maxWindowContigs, 2 * committedThisRun));
        } else if (cursor < items.size() && !items[cursor].fileEnd && !items[cursor].valid && committedThisRun >= curWindow) {
            curWindow = std::min<uint32_t>(cfg.maxWindowContigs, 2 * curWindow);  // a whole run committed: speculate further
        }
    }
}

void SequentialPass::loadRefPieces(int owner, const std::vector<ExtPiece>& pieces, bool rc) {
    std::vector<B200SparseEMMatcher::StagePiece> sp;
    sp.reserve(pieces.size());
    for (const auto& p : pieces) sp.push_back(B200SparseEMMatcher::StagePiece{p.dText, p.len, p.rc});
    rs.loadRefConcatFromReplica(owner, sp.data(), sp.size(), rc, cfg.refRegionSeparators, cfg.separator);
}

}  // namespace mbgc_b200
