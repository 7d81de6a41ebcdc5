/* include/mbgc_b200.h -- C ABI of the B200-native MBGC match finder (libmbgc_b200.so).
 *
 * This is the drop-in boundary for MBGC's one data-parallel hot path: the sparse k-mer exact-match
 * finder.  The reference has no FFI; its seam is the C++ class SlidingWindowSparseEMMatcher
 * (/root/reference/matching/SlidingWindowSparseEMMatcher.h:12-126, "ref .h" below; implementation
 * "ref .cpp" = matching/SlidingWindowSparseEMMatcher.cpp).  Every entry point here replaces one member of
 * that class, cited per declaration.  mbgc_b200/dropin/SlidingWindowSparseEMMatcher_b200.cpp re-implements
 * the class against the UNMODIFIED reference header on top of these calls (link-time substitution, see
 * INTEGRATION.md).
 *
 * Conventions
 *   - plain pointers and sizes only; all text is raw ASCII bytes (case-sensitive, N and IUPAC included --
 *     exactly what the reference hashes and compares);
 *   - every function that can fail returns 0 on success and a negative MBGCG_E_* code otherwise (the
 *     reference prints to stderr and exit()s; the shim converts); mbgcg_last_error() gives the text;
 *   - there is NO CPU fallback: without a CUDA device / the sm_100a kernels, mbgcg_create fails;
 *   - results are bit-exact with the reference run deterministically (-t 1 / OMP_NUM_THREADS=1);
 *   - a handle is not thread-safe; callers serialise calls on one handle (the reference's parallel mode
 *     enters matchTexts concurrently -- the shim takes a mutex).
 */
#ifndef MBGC_B200_H
#define MBGC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBGCG_NO_LOCK UINT64_MAX /* SW_END_ERASED_FLAG, ref .h:46 */

enum {
    MBGCG_OK = 0,
    MBGCG_E_CUDA = -1,     /* CUDA runtime / launch failure, or no device */
    MBGCG_E_PARAM = -2,    /* the reference's own parameter errors (e.g. minMatchLength < K, ref .cpp:480-483) */
    MBGCG_E_UNSUPPORTED = -3, /* outside the accelerated scope (k2 != 1, 64-bit table: ref .cpp:505-508) */
    MBGCG_E_LOCK = -4,     /* "Invalid worker lock value", ref .cpp:388-391 */
    MBGCG_E_NOMEM = -5
};

typedef struct mbgcg mbgcg_t;

/* Same layout as PgTools::TextMatch (matching/TextMatchers.h:9-17): 4 x uint64. */
typedef struct {
    uint64_t posSrcText;
    uint64_t length;
    uint64_t posDestText;
    uint64_t nextSrcRegionLoadingPos; /* always 0 on return, as in the TextMatch ctor */
} mbgcg_match_t;

/* Host-visible matcher state = the protected members the reference header's inline getters/setters touch
 * (ref .h:29-33,45-49,78,93-113). */
typedef struct {
    int64_t pos1;
    uint64_t maxRefLength;
    int32_t reachedRefLengthCount;
    int32_t circularBuffer;
    uint64_t swSize;
    uint64_t swEnd;
    uint64_t samplingPos;
    uint32_t hashSize;
    int32_t K;
} mbgcg_state_t;

/* ---- life cycle ----------------------------------------------------------------------------------- */

/* SlidingWindowSparseEMMatcher / SlidingWindowExpSparseEMMatcher constructors (ref .cpp:325-359,494-519);
 * the variant is chosen as MultipleGenomeMatchingProcessor.cpp:170-176 does (ctz(k1) != 0 -> Exp).
 * Allocates on `device`: the reference buffer (refLengthLimit bytes), the 2^24..2^31-slot table and its
 * occupancy bitmap; plus a host mirror of the reference buffer for getRef(). */
int mbgcg_create(mbgcg_t** out, uint64_t refLengthLimit, uint32_t targetMatchLength, int k1, int k2, int skipMargin,
                 int device);
/* The same matcher replicated on `ndev` devices: every replica holds the whole reference buffer and
 * index; every state-changing call below (window configuration, load_ref, load_separator, locks, set_state,
 * set_position) is applied to all of them in the same order, so the replicas stay bit-identical without any table
 * contents travelling.  Single-text calls (mbgcg_match, mbgcg_match_batch*, mbgcg_pack_batch, mbgcg_ingest_fasta) and
 * all getters use replica 0; the collection pass (mbgcg_seq_*) shards its windows over all replicas.
 * mbgcg_create(..., device) == mbgcg_create_multi(..., &device, 1). */
int mbgcg_create_multi(mbgcg_t** out, uint64_t refLengthLimit, uint32_t targetMatchLength, int k1, int k2, int skipMargin,
                       const int* devices, int ndev);
int mbgcg_device_count(const mbgcg_t* h);
/* ~SlidingWindowSparseEMMatcher (ref .cpp:460-467) */
void mbgcg_destroy(mbgcg_t* h);
const char* mbgcg_last_error(const mbgcg_t* h);
/* text of the last error of a failed mbgcg_create (no handle exists then) */
const char* mbgcg_create_error(void);

/* ---- window configuration (ref .h:93-97) ------------------------------------------------------------ */
void mbgcg_disable_sliding_window(mbgcg_t* h);
void mbgcg_disable_circular_buffer(mbgcg_t* h);
void mbgcg_set_sliding_window_size(mbgcg_t* h, int factor);

/* ---- reference loading ------------------------------------------------------------------------------ */

/* loadRef(refText, refLength, loadRCRef, addRegionSeparators, regionSeparator)  (ref .cpp:453-458 -> 402-437
 * -> index build 146-171).  text is a HOST pointer. */
int mbgcg_load_ref(mbgcg_t* h, const char* text, uint64_t len, int loadRC, int addSeparators, char separator);
/* loadSeparator (ref .cpp:439-451) */
int mbgcg_load_separator(mbgcg_t* h, char separator);

/* ---- matching ------------------------------------------------------------------------------------------ */

/* matchTexts(resMatches, destText, destLen, false, false, minMatchLength, matchingLockPos)
 * (ref .cpp:477-492 -> 181-321).  text is a HOST pointer.  *out points at handle-owned storage valid until
 * the next match call (the reference clears the caller's vector first, ref .cpp:484). */
int mbgcg_match(mbgcg_t* h, const char* text, uint64_t len, uint32_t minMatchLength, uint64_t lockPos,
                const mbgcg_match_t** out, uint64_t* n);

/* n independent matchTexts calls against the SAME frozen matcher state, in one device pass (the unit the
 * multi-GPU sharding and the speculative sequential schedule work in).  lockPos may be NULL (no locks).
 * out: all lists concatenated in text order; counts[i] = matches of text i.
 * out == NULL (here and in mbgcg_match_batch_device): the lists stay on the device -- only the counts come back --
 * for an mbgcg_pack_batch(matches = NULL) call right after; a caller that only wants the packed streams never pays
 * the 32 bytes per match of the device->host copy. */
int mbgcg_match_batch(mbgcg_t* h, int nTexts, const char* const* texts, const uint64_t* lens, uint32_t minMatchLength,
                      const uint64_t* lockPos, const mbgcg_match_t** out, const uint64_t** counts);

/* Same, but the texts are already resident in device memory (one buffer; text i occupies
 * [offsets[i], offsets[i]+lens[i]); offsets must be multiples of 16 with >= 64 zero bytes between texts and
 * >= 8192 readable bytes after the last one).  Nothing is copied host<->device except the final match
 * lists.  stream: a cudaStream_t (or NULL for the handle's own stream). */
int mbgcg_match_batch_device(mbgcg_t* h, const void* dText, int nTexts, const uint64_t* offsets, const uint64_t* lens,
                             uint32_t minMatchLength, const uint64_t* lockPos, const mbgcg_match_t** out,
                             const uint64_t** counts);

/* ---- stream packing, K6 -------------------------------------------------------------------------------------
 * Replaces MBGC_Encoder::processMatches with extendMatchLeft / extendMatchRight
 * (/root/reference/mbgccoder/MBGC_Encoder.cpp:143-308, 373-427, 310-371) and
 * ContextAwareMismatchesCoder::mismatch2code (coders/ContextAwareMismatchesCoder.cpp:22-27): the match list of a
 * text becomes the bytes the reference appends to targetLiterals / targetMapOffDests / targetMapOff5thByte /
 * targetMapLenDests / targetGapDeltas / targetGapMismatchesFlags, plus the list as pass 1 leaves it and the returned
 * unmatched count.  The caller keeps what needs encoder state: the dissimilar-contig bail-out (:202-205, decided on
 * the unmatched count returned here) and the statistics counters; processLiteral (:128-135, off by default) is
 * mbgcg_set_literal_ref_ext / mbgcg_pack_literal_ref_ext below. */

/* the MBGC_Params members processMatches reads (mbgccoder/MBGC_Params.h:66-97); field names as in the reference */
typedef struct {
    int32_t enableExtensionsWithMismatches;
    int32_t gapDepthOffsetEncoding;       /* 0..127 here (the reference allows 128) */
    int32_t gapDepthMismatchesEncoding;   /* 0..2 here (larger values are a developer-build option of the reference) */
    int32_t lazyDecompressionSupport;
    int32_t frugal64bitLenEncoding;
    int32_t enable40bitReference;
    int32_t mismatchesWithExclusion;
    int32_t mmsMismatchesInitialScore;
    int32_t mmsMismatchesScoreThreshold;
    int32_t mmsMismatchPenalty;
    int32_t mmsMatchBonus;
    int32_t reserved;
    uint64_t gapBreakingMatchMinLength;
} mbgcg_pack_params_t;

enum {
    MBGCG_S_LITERALS = 0,  /* targetLiterals: literals, mismatch codes, MATCH_MARK (MBGC_Params.h:45) */
    MBGCG_S_MAPOFF = 1,    /* targetMapOffDests: u32 LE posSrcText of every unpaired match */
    MBGCG_S_MAPOFF5 = 2,   /* targetMapOff5thByte (enable40bitReference) */
    MBGCG_S_MAPLEN = 3,    /* targetMapLenDests: frugal u16/u32/u64 (utils/helper.cpp:237-246) or plain u32 */
    MBGCG_S_GAPDELTA = 4,  /* targetGapDeltas */
    MBGCG_S_GAPFLAGS = 5,  /* targetGapMismatchesFlags */
    MBGCG_S_COUNT = 6
};

typedef struct {
    const uint8_t* stream[MBGCG_S_COUNT];      /* all texts concatenated in text order; handle-owned, valid until the next pack call */
    const uint64_t* streamOff[MBGCG_S_COUNT];  /* [nTexts+1]: text i owns stream[s][streamOff[s][i] .. streamOff[s][i+1]) */
    const uint64_t* unmatched;                 /* [nTexts] processMatches' return value (unmatched chars before extension) */
    const uint64_t* kept;                      /* [nTexts] matches left after pass 1 (gap-breaking matches removed) */
} mbgcg_pack_result_t;

/* processMatches(textMatches, destStart, destLen, targetIdx, matchingLockPos) for n texts in one device pass.
 * texts / matches are HOST pointers; matches = all lists concatenated, counts[i] of them belong to text i.
 * matches == NULL: pack the lists of the immediately preceding mbgcg_match / mbgcg_match_batch[_device] call on this
 * handle (same texts, single pass) -- lists and texts are still on the device, nothing is uploaded again (texts may
 * then be NULL too).  lockPos may be NULL (all SIZE_MAX).  loadedPosArr = refExtLoadedPosArr (MBGC_Encoder.h:47), read by
 * the lazy-decompression pairing rule (:252-261). */
int mbgcg_pack_batch(mbgcg_t* h, const mbgcg_pack_params_t* p, int nTexts, const char* const* texts, const uint64_t* lens,
                     const mbgcg_match_t* matches, const uint64_t* counts, const uint64_t* lockPos,
                     const uint64_t* loadedPosArr, uint64_t nLoadedPos, mbgcg_pack_result_t* out);
/* textMatches as pass 1 of the last pack call left them (MBGC_Encoder.cpp:197), all texts concatenated (kept[] each) */
int mbgcg_pack_kept_matches(mbgcg_t* h, const mbgcg_match_t** out, uint64_t* n);
/* MBGC_Encoder::processLiteral (/root/reference/mbgccoder/MBGC_Encoder.cpp:128-135, called from processMatches at :228
 * and :290; params: MGMP_Params.h:68-69 refLiteralMinimalLengthExt / refLiteralBeforeAfterExt, :175-177
 * isLiteralProperForRefExtension; a developer-build option of the reference, `-N` / `-F` of main.cpp:308-313, off by
 * default).  After mbgcg_set_literal_ref_ext every pack call on the handle also lists what the reference would append to
 * targetRefExtensions[targetIdx]: for a text that is NOT proper for a whole-contig reference extension (:211,
 * unmatched * unmatchedFractionFactor <= length, MGMP_Params.h:179-187) every literal run longer than the threshold, with
 * up to refLiteralBeforeAfterExt bytes of context on either side, as a range of that text, in append order.  The
 * extension string of text i = the concatenation of text_i[start, start + len) over its ranges.
 * refLiteralMinimalLengthExt = UINT64_MAX switches the option off again. */
typedef struct {
    uint32_t text;      /* index into the texts of the pack call */
    uint32_t start;     /* pos - before */
    uint32_t len;       /* length + before + after */
    uint32_t reserved;
} mbgcg_lit_range_t;
int mbgcg_set_literal_ref_ext(mbgcg_t* h, uint64_t refLiteralMinimalLengthExt, uint64_t refLiteralBeforeAfterExt,
                              int unmatchedFractionFactor);
/* the ranges of the last pack call (handle-owned, valid until the next pack call) */
int mbgcg_pack_literal_ref_ext(mbgcg_t* h, const mbgcg_lit_range_t** ranges, uint64_t* n);
/* last pack call: ms[0..3] = input transfer, kernels, stream D2H, total (CUDA events); kernel launches; fixed-point
 * sweeps used for (removal, pairing) */
void mbgcg_pack_timing(const mbgcg_t* h, float ms[4], uint64_t* launches, int sweeps[2]);

/* ---- FASTA ingest ---------------------------------------------------------------------------------------------
 * Replaces the loop over the reference's FASTA reader for one file: kseq_read_lossless_fasta
 * (/root/reference/utils/kseq.h:233-274; called at matching/MultipleGenomeMatchingProcessor.cpp:9-14, 126-139,
 * 400-406): record splitting at every '>' that starts a line, the header line, the sequence with its line ends
 * removed, the DNA line length detection with its "not well-formed" exits.  The file image goes to the device as it is
 * on disk (gzip stays with libdeflate on the host); the sequences stay there as a batch of texts. */
typedef struct {
    uint64_t hdrOff, hdrLen;   /* seq->name: the header line inside `bytes`, without '>' and without the '\n' */
    uint64_t textOff, seqLen;  /* seq->seq: the sequence inside the device buffer returned in *dTexts */
} mbgcg_fasta_rec_t;

#define MBGCG_DNA_NOT_WELLFORMED UINT64_MAX /* DNA_NOT_WELLFORMED in the reader's uint64 member, utils/kseq.h:35,340 */

/* bytes = the whole (decompressed) file, HOST pointer.  *recs / *nRecs: the records every successful call of the reader
 * would have returned, in file order (handle-owned, valid until the next ingest).  *status: the negative value that ends
 * the reader's loop (-1 end of file, -3 not FASTA, -4 DNA not well-formed -- MBGC exits on the last two,
 * MultipleGenomeMatchingProcessor.cpp:17-30).  *dnaLineLen: seq->dnaLineLen after the last call (what KSEQ_DNA_LINE_LENGTH
 * reads).  *dTexts: device buffer holding the sequences at textOff (multiples of 16, >= 64 zero bytes between them,
 * 8 KiB behind the last): pass it with textOff / seqLen to mbgcg_match_batch_device. */
int mbgcg_ingest_fasta(mbgcg_t* h, const char* bytes, uint64_t n, const mbgcg_fasta_rec_t** recs, uint64_t* nRecs,
                       const void** dTexts, uint64_t* dnaLineLen, int* status);
/* copies sequence `record` of the last ingest to host memory (seqLen bytes) */
int mbgcg_copy_ingested(mbgcg_t* h, uint64_t record, char* out);
/* last ingest: ms[0..2] = file image host -> device, kernels (incl. two small host round trips), total (CUDA events);
 * kernel launches */
void mbgcg_ingest_timing(const mbgcg_t* h, float ms[3], uint64_t* launches);

/* ---- the collection pass: row a9 of SURVEY.md 8 -----------------------------------------------------------------
 * Replaces the target loop of the reference's sequential schedule,
 * MultipleGenomeMatchingProcessor::processTargetsWithParallelIO
 * (/root/reference/matching/MultipleGenomeMatchingProcessor.cpp:232-313: per contig matchTexts :274 -> processMatches :276
 * -> isContigProperForRefExtension / ...RCExtension :277-285 (MGMP_Params.h:179-196) -> loadRef :286; per file
 * processTargetMeta :297 and the hook processAfterTargetWithParallelIO, mbgccoder/MBGC_Encoder.cpp:493-509: loadSeparator,
 * refExtSize, acquire/releaseWorkerMatchingLockPos), batched and sharded over the handle's devices:
 *
 *   file images (as on disk, after gunzip) are pushed in target order; groups of them go to the replicas in turn, where
 *   the FASTA ingest, the match stage (lists stay on the device) and the stream packer run for a whole window of contigs
 *   against the frozen matcher state; results are committed in target order up to and including the first contig that
 *   extends the reference; the extension reaches every replica device-to-device and the rest of the window runs again.
 *
 * What comes out is byte-identical to the one-contig-at-a-time schedule (`mbgc -t 1`).  The caller keeps what belongs to
 * the encoder: headers, stream concatenation, separator marks, meta streams.  While a pass is in progress no other call
 * may be made on the handle; the host mirror behind mbgcg_host_ref is switched off by mbgcg_seq_begin (extensions come
 * from device memory). */
typedef struct {
    mbgcg_pack_params_t pack;
    uint32_t minMatchLength;            /* params->k (MultipleGenomeMatchingProcessor.cpp:274) */
    int32_t unmatchedFractionFactor;    /* currentUnmatchedFractionFactor (MGMP_Params.h:41,179-187) */
    int32_t unmatchedFractionRCFactor;  /* unmatchedFractionRCFactor (MGMP_Params.h:42,189-191) */
    int32_t rcInReferenceDisabled;      /* isRCinReferenceDisabled() (MGMP_Params.h:82-84) */
    int32_t refRegionSeparators;        /* params->refRegionSeparators */
    int32_t separatorPerFile;           /* lazyDecompressionSupport: loadSeparator + a refExtLoadedPosArr entry per file */
    char regionSeparator;               /* MGMP_Params::REF_REGION_SEPARATOR */
    uint64_t chunkBytes;                /* file-image bytes per ingest slot; 0 = default (64 MiB) */
    uint32_t maxWindowContigs;          /* speculation window; 0 = default */
    /* 1: the batch-frozen schedule of the reference's PARALLEL mode instead of the sequential one
     * (processTargetsParallel / processTarget / finalizeParallelProcessingOfTarget,
     * MultipleGenomeMatchingProcessor.cpp:340-468, 520-555): every target (file) of a window is matched under its own
     * sliding-window lock against the matcher state at the window's start; its qualifying contigs -- each followed by its
     * reverse complement when contigsIndividuallyReversed (:389-398) -- extend the reference as ONE text when the target is
     * finalized, in target order (:441).  No result is ever redone except for the dissimilar-contig rule
     * (MBGC_Encoder.cpp:202-205).  Any such schedule yields a legal archive; this one is deterministic.
     * Events: CONTIG carries contigLock; FILE_END carries the lock in lockPos and refExtSize measured across the
     * extension and the separator. */
    int32_t parallelSchedule;
    int32_t contigsIndividuallyReversed;                       /* MGMP_Params.h:75 */
    int32_t allowedTargetsOutrun;                              /* targets per window: matcherWorkingThreads * allowedTargetsOutrunFactor */
    int32_t allowedTargetsOutrunForDissimilarContigs;          /* MGMP_Params.h:78 */
    int32_t unmatchedFractionFactorTweakForDissimilarContigs;  /* MGMP_Params.h:80 */
    uint64_t minimalLengthForDissimilarContigs;                /* MGMP_Params.h:79 */
} mbgcg_seq_config_t;

enum { MBGCG_EV_CONTIG = 1, MBGCG_EV_FILE_END = 2, MBGCG_EV_NEED_INPUT = 3, MBGCG_EV_DONE = 4 };

typedef struct {
    int32_t type;
    int32_t status;          /* FILE_END: what ended the reader's loop for this file: -1 end of file, -3 not FASTA, -4 DNA
                              * not well-formed (MBGC exits on the last two, MultipleGenomeMatchingProcessor.cpp:17-30) */
    uint64_t file;           /* index of the file in push order */
    /* CONTIG: one committed contig, in target order */
    const char* hdr;         /* the header line (seq->name of the lossless reader), valid until the next call */
    uint64_t hdrLen;
    uint64_t seqLen;
    uint64_t nMatches;       /* resMatches.size() */
    uint64_t kept;           /* matches left after pass 1 of processMatches */
    uint64_t unmatched;      /* processMatches' return value */
    int32_t extended;        /* the contig was loaded into the reference (parallel schedule: will be, with its target) ... */
    int32_t extendedRC;      /* ... and its reverse complement */
    uint64_t contigLock;     /* parallel schedule: the matchingLockPos the contig was matched and packed under */
    uint64_t window;         /* running number of the window (group of files ingested together) the file belongs to */
    const uint8_t* seg[MBGCG_S_COUNT];   /* what processMatches appends to the six streams; valid until the next call */
    uint64_t segLen[MBGCG_S_COUNT];
    /* FILE_END */
    uint64_t nContigs;       /* seqCounter */
    uint64_t dnaLineLen;     /* KSEQ_DNA_LINE_LENGTH(seq) */
    uint64_t refExtSize;     /* getLoadedRefLength() - startPos after the separator (MBGC_Encoder.cpp:502) */
    uint64_t lockPos;        /* the value appended to locksPosStream (MBGC_Encoder.cpp:506-508) */
    uint64_t loadedRefLength;/* getLoadedRefLength() afterwards (also in DONE) */
} mbgcg_seq_event_t;

typedef struct {
    uint64_t files, contigs, windows, runs, rerunContigs, extensions;
    uint64_t imageBytes, textBytes, pageableBytes, launches, chunks;
    double ingestMs, matchMs, packMs, loadRefMs, waitStageMs;  /* host wall time inside the respective calls */
    uint64_t wireBytes;   /* image bytes that crossed PCIe (packed images travel at about a quarter of their size) */
} mbgcg_seq_stats_t;

/* loadedPosArr = refExtLoadedPosArr as it stands (MBGC_Encoder.cpp:786-791); the pass appends to its own copy */
int mbgcg_seq_begin(mbgcg_t* h, const mbgcg_seq_config_t* cfg, const uint64_t* loadedPosArr, uint64_t nLoadedPos);
/* next file image, HOST pointer.  Memory from mbgcg_image_alloc now belongs to the pass (it is released as soon as the
 * image is on the device); any other memory must stay valid until the file's FILE_END event has been returned. */
int mbgcg_seq_push_file(mbgcg_t* h, const char* bytes, uint64_t n);
/* The same for an image that was packed on the host (below): 4 bases per byte cross PCIe, the device rebuilds the file byte for
 * byte before the reader's kernels look at it. */
int mbgcg_seq_push_packed(mbgcg_t* h, const void* packedImage);
int mbgcg_seq_finish_input(mbgcg_t* h);
/* 1 while the pass can use more images to keep the devices busy (about two windows ahead of the commit cursor) */
int mbgcg_seq_wants_input(mbgcg_t* h);
/* next event in target order; MBGCG_EV_NEED_INPUT = nothing can be committed before another file (or finish_input) */
int mbgcg_seq_next(mbgcg_t* h, mbgcg_seq_event_t* ev);
int mbgcg_seq_end(mbgcg_t* h, mbgcg_seq_stats_t* stats);

/* The same pass run to its end in one call, for callers that want the encoder's view of the result rather than the
 * events: the six streams as MBGC_Encoder holds them after the last target -- every contig's segments in target order,
 * SEQ_SEPARATOR_MARK behind each contig's literals (processAfterSequence, MBGC_Encoder.cpp:489-491), FILE_SEPARATOR_MARK
 * behind each file's mismatch flags (processAfterTarget, :493-496) -- plus the per-file meta data and the header lines.
 * next_image is called from the calling thread whenever the pass can use another file: return 1 with *bytes / *n set,
 * 0 if none is ready yet (only allowed when block == 0), -1 when there are no more files. */
typedef int (*mbgcg_next_image_fn)(void* user, int block, const char** bytes, uint64_t* n);
typedef struct {
    const uint8_t* stream[MBGCG_S_COUNT];
    uint64_t streamLen[MBGCG_S_COUNT];
    uint64_t nFiles, nContigs;
    const uint64_t* seqCounts;     /* [nFiles] processTargetMeta's first argument */
    const uint64_t* dnaLineLens;   /* [nFiles] ... and its second */
    const uint64_t* refExtSizes;   /* [nFiles] what goes to refExtSizeDest (lazy decompression support) */
    const uint64_t* lockPos;       /* [nFiles] what goes to locksPosStream */
    const int32_t* status;         /* [nFiles] the reader's final status per file */
    const char* headers;           /* all header lines, each followed by '\n', in target order */
    uint64_t headersLen;
    uint64_t matches, unmatched, extensions;
    uint64_t digest;               /* sum over the contigs of a 64-bit FNV-1a hash of their stream segments: independent
                                    * of how the files were split over windows, devices or ranks */
} mbgcg_seq_result_t;
/* call between mbgcg_seq_begin and mbgcg_seq_end; *out stays valid until mbgcg_seq_end */
int mbgcg_seq_run(mbgcg_t* h, mbgcg_next_image_fn next_image, void* user, int keepStreams, mbgcg_seq_result_t* out);

/* mbgcg_seq_run for files on disk: `readers` threads inside the library bring them into pinned memory ahead of the pass (in
 * order of need; read() into the arena, or packed 4 bases per byte on the way when packImages != 0 -- mbgcg_image_pack_file),
 * files that do not fit the arena through pageable memory.  gzip is the caller's business (libdeflate stays on the host side of
 * the reference).  hostMs (may be NULL): [0] time the readers spent reading / packing, summed over threads, [1] time the pass
 * waited for a file. */
int mbgcg_seq_run_files(mbgcg_t* h, const char* const* paths, uint64_t nFiles, int readers, int packImages, int keepStreams,
                        mbgcg_seq_result_t* out, double hostMs[2]);

/* Pinned host memory for file images, cut from large slabs (cudaHostAlloc costs ~0.5 ms per MiB): DMA reads it at PCIe
 * speed, pageable images travel through bounce buffers at ~1/4 of that.  NULL when the arena is full -- use any other
 * memory then.  mbgcg_image_free only for images that were never pushed. */
void* mbgcg_image_alloc(mbgcg_t* h, uint64_t n);
/* A file image packed for the trip to the device: header, 2-bit codes, exception runs for every byte that is not A/C/G/T (the
 * line ends of a FASTA file are ONE run: first offset, count, distance).  Reading the file from the page cache and writing a
 * quarter of it into pinned memory costs the host half the memory traffic of copying it there, and PCIe a quarter -- on a box
 * whose cores feed several GPUs that is what the whole pipeline is bound by.  dst: 16-byte aligned memory of
 * mbgcg_image_pack_bound(n) bytes (normally from mbgcg_image_alloc); *used = the bytes that have to travel.  Thread-safe; the
 * reader threads of a caller run it in parallel.  MBGCG_E_UNSUPPORTED: too many exceptional bytes (soft-masked or protein
 * text) -- push the file as it is.  src may be pageable (an mmap of the file).
 * mbgcg_seq_run's image source returns such an image with MBGCG_IMAGE_PACKED set in *n. */
#define MBGCG_IMAGE_PACKED (1ull << 63)
uint64_t mbgcg_image_pack_bound(uint64_t n);
int mbgcg_image_pack(const void* src, uint64_t n, void* dst, uint64_t dstCap, uint64_t* used);
/* the same for a file on disk: read() in pieces that stay in the core's cache and packed from there, so the page cache is read
 * once and a quarter of it written (sixteen threads faulting in mmaps of their files would queue up on the memory-map lock) */
int mbgcg_image_pack_file(const char* path, void* dst, uint64_t dstCap, uint64_t* used, uint64_t* fileSize);
/* the largest image the arena can hold in one piece (larger requests always return NULL) */
uint64_t mbgcg_image_max(mbgcg_t* h);
int mbgcg_image_free(mbgcg_t* h, const void* p);

/* switches the host mirror behind mbgcg_host_ref off / on (only the reference's own CPU packer reads it,
 * MBGC_Encoder.cpp:181,318-330,376-380); while it is off mbgcg_load_ref skips the host-side copy */
void mbgcg_set_host_mirror(mbgcg_t* h, int on);

/* ---- the reverse-complement redundancy pass of mode 3 (row f3 of SURVEY.md 8) ------------------------------------
 * SimpleSequenceMatcher::rcMatchSequence(sequence, rcMapOff, rcMapLen, targetMatchLength, minMatchLength)
 * (/root/reference/matching/SimpleSequenceMatcher.cpp:165-176; called on the literal stream at
 * mbgccoder/MBGC_Encoder.cpp:637-638) over CopMEMMatcher (matching/copmem/CopMEMMatcher.cpp) with the single-threaded
 * (`-t 1`) index build: stretches of `seq` that are the reverse complement of an earlier stretch are cut out and replaced by
 * RC_MATCH_MARK; mapOff gets their source positions (u32 each), mapLen the frugal-byte lengths (first entry: the minimal
 * match length).  seq (HOST, n bytes) is rewritten in place, *newLen = its new length.  mapOff needs 4 and mapLen 10 bytes
 * per (n / minMatchLength + 2) matches.  minMatchLength = UINT32_MAX means targetMatchLength, as in the reference.
 * deviceMs (may be NULL): [0] index build, [1] query, CUDA events. */
int mbgcg_rc_match_sequence(mbgcg_t* h, uint8_t* seq, uint64_t n, uint32_t targetMatchLength, uint32_t minMatchLength, uint64_t* newLen,
                            uint8_t* mapOff, uint64_t* mapOffLen, uint8_t* mapLen, uint64_t* mapLenLen, float deviceMs[2]);

/* ---- the decoder's match-copy path (row f4 of SURVEY.md 8) -------------------------------------------------------------
 * What MBGC_Decoder does between "the streams are decompressed" and "the DNA of every file is in memory"
 * (/root/reference/mbgccoder/MBGC_Decoder.cpp: initReference :243-262 or decodeReference :264-317 for G0, then per target
 * decodeTarget :532-649 = per contig decodeSequenceAndReturnUnmatchedChars :319-432 with extendMatchRight / extendMatchLeft
 * :434-527, the extension decision of matching/MGMP_Params.h:179-190 and loadRef :651-675).  Headers, line breaks, file output,
 * lazy / filtered extraction and the back-end coders stay with the caller.  All targets are parsed at once on the device and
 * written in as few launches as the growth of the reference allows. */
typedef struct {
    const uint8_t* stream[MBGCG_S_COUNT];  /* HOST: literalStr (after restoreRCMatchedSequence in mode 3), mapOffStream,
                                            * mapOff5thByteStream, mapLenStream (as archived), gapDeltaStr, gapMismatchesFlagsStr */
    uint64_t streamLen[MBGCG_S_COUNT];
    uint64_t refTotalLength;               /* readStats :1309-1311 */
    int32_t g0IsTarget;                    /* params->g0IsTarget: initReference (the first contig seeds the reference, file 0 is a
                                            * target) instead of decodeReference (G0 = g0SeqCount verbatim contigs) */
    uint32_t g0SeqCount;
    int32_t rcInReference;                 /* !isRCinReferenceDisabled() */
    int32_t contigsIndividuallyReversed;   /* params->contigsIndivduallyReversed */
    uint32_t nTargets;
    const uint32_t* seqCounts;             /* [nTargets] seqsCountArr of the targets */
    const uint8_t* unmatchedFractionFactors; /* [2 * nTargets] as archived (factor, RC factor) */
    const uint64_t* lockPos;               /* [nTargets] matchingLocksPosArr, or NULL (= refTotalLength, :587) */
} mbgcg_decode_input_t;
typedef struct {
    uint64_t nContigs;                 /* decoded contigs in output order: G0's (decodeReference only), then every target's */
    const uint64_t* contigLen;         /* handle-owned, valid until the next decode call */
    const uint64_t* contigUnmatched;   /* unmatchedChars of decodeSequenceAndReturnUnmatchedChars (0 for verbatim G0 contigs) */
    uint64_t dnaLen;                   /* sum of contigLen */
    uint64_t refPos;                   /* the decoder's refPos and reachedRefLengthCount after the last target */
    int32_t reachedRefLengthCount;
    uint32_t batches;                  /* launches of the copy pass */
    float parseMs, copyMs;             /* CUDA events: uploads + markers + parse; copy pass + reference growth */
    uint64_t launches;
} mbgcg_decode_result_t;
/* p: the fields the decoder reads (extensions, gap depths, lazy support, frugal lengths, the four score values) */
int mbgcg_decode_targets(mbgcg_t* h, const mbgcg_pack_params_t* p, const mbgcg_decode_input_t* in, mbgcg_decode_result_t* out);
/* contigs [first, first + n) of the last decode call, back to back, to HOST memory (pinned memory travels at PCIe speed) */
int mbgcg_decode_fetch(mbgcg_t* h, uint64_t firstContig, uint64_t n, uint8_t* dst);
/* the same contigs where they are: a device pointer to contig `firstContig` (valid until the next decode call) */
const void* mbgcg_decode_device_dna(mbgcg_t* h, uint64_t firstContig);
/* bytes [pos, pos + n) of the decoder's reference buffer (refStr) to HOST memory */
int mbgcg_decode_fetch_ref(mbgcg_t* h, uint64_t pos, uint64_t n, uint8_t* dst);

/* ---- the two legacy symbol packers the reference carries (row a14 of SURVEY.md 8; no live call site in MBGC) ----
 * SymbolsPackingFacility::packSequence (/root/reference/coders/SymbolsPackingFacility.cpp:123-139, packSuffixSymbols
 * :149-159): symbols of an alphabet of nSymbols letters (`symbols`, in order), as many per byte as fit -- 4 for ACGT,
 * 3 for ACGNT -- first symbol most significant, a short tail padded with zeros.  dst needs ceil(n / perByte) bytes.
 * A byte that is not in the alphabet is an error (the reference's validateSymbol ends the program). */
int mbgcg_symbols_pack(mbgcg_t* h, const char* symbols, int nSymbols, const char* src, uint64_t n, uint8_t* dst, uint64_t* nOut);
/* SymbolsPackingFacility::reverseSequence(sequence, pos, length, destPtr) (:190-215): symbols pos .. pos+length-1 */
int mbgcg_symbols_unpack(mbgcg_t* h, const char* symbols, int nSymbols, const uint8_t* packed, uint64_t pos, uint64_t length, char* dst);
/* VarLenDNACoder::encode (/root/reference/coders/VarLenDNACoder.cpp:55-112) for a code book given as the reference reads
 * it (initUsing :10-36: up to 256 words of 1..4 symbols separated by '\n', the first a single symbol; writeBook :41-52
 * produces it): greedy longest-first parse, one output byte per word.  The three static books of the reference are data
 * of the reference; pass any of them, or your own.  n < 4 takes the tail rule of :89-111 only (the reference's main loop
 * underflows there). */
int mbgcg_varlen_encode(mbgcg_t* h, const char* codeBook, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dstCap, uint64_t* nOut);
/* VarLenDNACoder::decode (:114-128): every code byte expands to its word; fails unless exactly expLen symbols come out */
int mbgcg_varlen_decode(mbgcg_t* h, const char* codeBook, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t expLen);

/* ---- sliding-window locks (ref .cpp:361-400) ------------------------------------------------------------ */
uint64_t mbgcg_acquire_lock(mbgcg_t* h);
int mbgcg_release_lock(mbgcg_t* h, uint64_t lockValue);

/* ---- getters / setters (ref .h:104-113) ------------------------------------------------------------------ */
const char* mbgcg_host_ref(const mbgcg_t* h);            /* getRef(): host mirror of the reference buffer */
uint64_t mbgcg_max_ref_length(const mbgcg_t* h);         /* getMaxRefLength() */
uint64_t mbgcg_ref_length(const mbgcg_t* h);             /* getRefLength() */
uint64_t mbgcg_loading_position(const mbgcg_t* h);       /* getLoadingPosition() */
uint64_t mbgcg_loaded_ref_length(const mbgcg_t* h);      /* getLoadedRefLength() */
void mbgcg_set_position(mbgcg_t* h, uint64_t refPos, int reachedRefLengthCount); /* setPosition() */
void mbgcg_get_state(const mbgcg_t* h, mbgcg_state_t* out);
/* push the members the reference header's inline methods may have changed behind our back */
void mbgcg_set_state(mbgcg_t* h, const mbgcg_state_t* in);

/* ---- inspection (tests / bench) ------------------------------------------------------------------------- */

/* Copies the low 32 bits of every table slot (= the reference's ht32bit content) to host memory. */
int mbgcg_copy_table(mbgcg_t* h, uint32_t* out, uint64_t nSlots);
/* Copies the device reference buffer [from, from+n) to host memory (to check it equals the mirror). */
int mbgcg_copy_device_ref(mbgcg_t* h, uint64_t from, uint64_t n, char* out);

/* Copies text `textIndex` of the last mbgcg_match / mbgcg_match_batch call back from the device, as the kernels
 * saw it (host texts may travel 2-bit packed and are expanded on the device: this checks they arrive intact).
 * Only meaningful for calls that fit one pass. */
int mbgcg_copy_staged_text(mbgcg_t* h, uint64_t textIndex, char* out);

/* Host-only self-test of the 2-bit text packer (no device needed): packs text[0..n) as the transfer path does and
 * expands it again on the host into out[0..n).  Returns 0, or 1 if the text has too many non-ACGT bytes to be
 * packed (such a block travels raw; out is then untouched). */
int mbgcg_selftest_pack_image(const char* image, uint64_t n, char* out, uint64_t* packedBytes, uint64_t* nRuns);
int mbgcg_selftest_pack(const char* text, uint64_t n, char* out);

/* Host-only self-test of the reverse complement mbgcg_load_ref writes for loadRC (no device needed):
 * out[n - 1 - i] = complement(text[i]) with the table of /root/reference/utils/helper.cpp:312-358 (upper-casing, IUPAC,
 * everything else unchanged, 0x7F -> 0), i.e. upperReverseComplement (:405-410). */
void mbgcg_selftest_revcomp(const char* text, uint64_t n, char* out);

/* Stage timings of the last match call, CUDA events on the handle's stream, milliseconds. */
enum {
    MBGCG_T_H2D = 0,      /* host->device copy of the texts */
    MBGCG_T_PROBE = 1,    /* K3: hash + occupancy + table gather + K-mer verify -> seeds */
    MBGCG_T_COMPACT = 2,  /* scans + seed compaction + run detection */
    MBGCG_T_EXTEND = 3,   /* K4: bidirectional extension of seed runs */
    MBGCG_T_RESOLVE = 4,  /* K5: sequential-equivalent overlap/skip resolve -> match lists */
    MBGCG_T_D2H = 5,      /* device->host copy of the match lists */
    MBGCG_T_TOTAL = 6,    /* first to last event */
    /* parts of MBGCG_T_PROBE */
    MBGCG_T_ANCHOR = 7,   /* fast path: first 16 positions of every 512-position chunk */
    MBGCG_T_CLEAN = 8,    /* fast path: chunk == reference on the anchor's diagonal? */
    MBGCG_T_FULLPROBE = 9,/* full probe of the chunks the fast path left over (or of all chunks) */
    MBGCG_T_MEMO = 10,    /* self-probe of the reference (only when the index changed since the last one) */
    MBGCG_T_COUNT = 12
};
void mbgcg_last_timing(const mbgcg_t* h, float ms[MBGCG_T_COUNT]);
/* on = 0: match calls stop recording their stage events (mbgcg_last_timing then returns zeros): a dozen event records
 * less per call on the stream and on the calling thread.  Default: on. */
void mbgcg_set_stage_timing(mbgcg_t* h, int on);
/* counters of the last match call: 0 positions, 1 seed segments, 2 seed runs, 3 matches, 4 kernel launches,
 * 5 passes, 6 chunks (512 positions), 7 lanes (16 positions) that went through the full probe */
void mbgcg_last_counters(const mbgcg_t* h, uint64_t out[8]);
/* last match call: bytes that crossed PCIe, out[0] host -> device (texts raw or 2-bit packed, descriptors), out[1]
 * device -> host (counts + match lists) */
void mbgcg_last_wire_bytes(const mbgcg_t* h, uint64_t out[2]);
/* last mbgcg_load_ref: ms[0] device time of the call (staging copy, segment kernels, index inserts; CUDA events),
 * ms[1] host wall time; *appended = reference bytes it appended (text + reverse complement + separators) */
void mbgcg_load_ref_timing(const mbgcg_t* h, float ms[2], uint64_t* appended);
/* total kernel launches issued by this handle since creation */
uint64_t mbgcg_total_launches(const mbgcg_t* h);
/* the cudaStream_t every kernel and copy of this handle is issued on (to bracket calls with CUDA events) */
void* mbgcg_stream(const mbgcg_t* h);

#ifdef __cplusplus
}
#endif
#endif /* MBGC_B200_H */
