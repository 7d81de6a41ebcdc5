/* oracle/mbgc_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the MBGC hot path (the sparse k-mer exact-match finder
 * and the match/literal stream packing).  It exists to CHECK the CUDA path; nothing under oracle/ is
 * ever linked into, imported by, or executed from the product library (mbgc_b200/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every entry point below against the
 * reference's own code compiled from /root/reference (oracle/_ref/libref_matcher.so, oracle/_ref/mbgc)
 * on the reference's shipped fixtures (example-scripts/, 3 Listeria genomes) and seeded synthetic
 * collections, and tests/golden/ holds vectors generated from that reference build
 * (tests/golden/make_golden.py).
 *
 * Each function cites the reference file:line it restates (paths relative to /root/reference).
 */
#ifndef MBGC_ORACLE_H
#define MBGC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBO_NO_LOCK UINT64_MAX /* SW_END_ERASED_FLAG, matching/SlidingWindowSparseEMMatcher.h:46 */

typedef struct mbo_matcher mbo_matcher;

/* (posSrcText, length, posDestText) -- matching/TextMatchers.h:9-17 */
typedef struct {
    uint64_t src;
    uint64_t len;
    uint64_t dst;
} mbo_match_t;

/* utils/Hashes.h:28-40 : maRushPrime1HashSimplified<K>, low 32 bits */
uint32_t mbo_hash(const uint8_t* p, int K);
/* matching/SlidingWindowSparseEMMatcher.cpp:74-87 : hashed prefix length K from L and minMatchLength */
int mbo_kmer_from_L(uint32_t L, uint32_t minMatchLength);
/* matching/SlidingWindowSparseEMMatcher.cpp:99-103 : table size (power of two in [2^24, 2^31]) */
uint32_t mbo_hash_size(uint64_t maxRefLength, int k1);
/* utils/helper.cpp:312-358,405-410 : upper-casing reverse complement of src[0..n) into dst[0..n) */
void mbo_upper_reverse_complement(const uint8_t* src, uint64_t n, uint8_t* dst);
/* matching/SlidingWindowSparseEMMatcher.cpp:146-171 : which positions one index pass samples.
 * Writes up to cap (begin,end) pairs (step k1 inside each) to ranges, returns the number of pairs and the
 * new sampling position through *newSamplingPos. */
int mbo_insert_ranges(uint64_t samplingPos, int64_t refPos1, int K, int k1, int64_t* ranges, int cap,
                      uint64_t* newSamplingPos);

/* matching/SlidingWindowSparseEMMatcher.cpp:325-359,494-519 (+ MultipleGenomeMatchingProcessor.cpp:170-176:
 * k1 even -> "Exp" variant storing pos >> ctz(k1)) */
mbo_matcher* mbo_create(uint64_t refLengthLimit, uint32_t L, int k1, int k2, int skipMargin);
void mbo_destroy(mbo_matcher* m);
/* matching/SlidingWindowSparseEMMatcher.h:93-97 */
void mbo_disable_sliding_window(mbo_matcher* m);
void mbo_disable_circular_buffer(mbo_matcher* m);
void mbo_set_sliding_window_size(mbo_matcher* m, int factor);
/* matching/SlidingWindowSparseEMMatcher.cpp:402-437,453-458 */
void mbo_load_ref(mbo_matcher* m, const uint8_t* text, uint64_t len, int loadRC, int addSep, uint8_t sep);
/* matching/SlidingWindowSparseEMMatcher.cpp:439-451 */
void mbo_load_separator(mbo_matcher* m, uint8_t sep);
/* matching/SlidingWindowSparseEMMatcher.cpp:477-492 -> 181-321.  Returns the number of matches; the list
 * stays owned by the matcher until the next call (mbo_matches).  usePrefilter!=0 keeps the l/r 4-byte
 * side-probe filter of .cpp:203-204,223-226 (the reference's behaviour); 0 drops it (to test that it is
 * redundant). Returns UINT64_MAX on the reference's error path (minMatchLength < K). */
uint64_t mbo_match(mbo_matcher* m, const uint8_t* text, uint64_t len, uint32_t minMatchLength, uint64_t lockPos,
                   int usePrefilter);
const mbo_match_t* mbo_matches(const mbo_matcher* m);
/* matching/SlidingWindowSparseEMMatcher.cpp:361-400 */
uint64_t mbo_acquire_lock(mbo_matcher* m);
int mbo_release_lock(mbo_matcher* m, uint64_t lockValue);
/* matching/SlidingWindowSparseEMMatcher.h:104-113 */
const uint8_t* mbo_ref(const mbo_matcher* m);
uint64_t mbo_ref_length(const mbo_matcher* m);
uint64_t mbo_loading_position(const mbo_matcher* m);
uint64_t mbo_loaded_ref_length(const mbo_matcher* m);
uint64_t mbo_max_ref_length(const mbo_matcher* m);
void mbo_set_position(mbo_matcher* m, uint64_t refPos, int reachedCount);
const uint32_t* mbo_table(const mbo_matcher* m);
uint64_t mbo_table_size(const mbo_matcher* m);
uint64_t mbo_sampling_pos(const mbo_matcher* m);
int mbo_kmer(const mbo_matcher* m);
/* counters of the last mbo_match call: probes, non-empty slots, K-mer-equal candidates, accepted pushes */
void mbo_last_stats(const mbo_matcher* m, uint64_t out4[4]);

/* ---- stream packing: mbgccoder/MBGC_Encoder.cpp:143-427 -------------------------------------------- */

typedef struct {
    int enableExtensionsWithMismatches; /* MBGC_Params.h: mode presets */
    int gapBreakingMatchMinLength;
    int gapDepthOffsetEncoding;
    int gapDepthMismatchesEncoding;
    int lazyDecompressionSupport;
    int frugal64bitLenEncoding;
    int enable40bitReference;
    int mismatchesWithExclusion;
    int mmsMismatchesInitialScore;
    int mmsMismatchesScoreThreshold;
    int mmsMismatchPenalty;
    int mmsMatchBonus;
    int k;                               /* params->k (minimal match length) */
    int unmatchedFractionFactor;         /* for the ref-extension decision / loadLiteralsToRef flag */
} mbo_pack_params_t;

typedef struct mbo_streams mbo_streams;
mbo_streams* mbo_streams_create(void);
void mbo_streams_destroy(mbo_streams* s);
/* stream ids */
enum { MBO_S_LITERALS = 0, MBO_S_MAPOFF = 1, MBO_S_MAPOFF5 = 2, MBO_S_MAPLEN = 3, MBO_S_GAPDELTA = 4,
       MBO_S_GAPFLAGS = 5, MBO_S_COUNT = 6 };
const uint8_t* mbo_stream_data(const mbo_streams* s, int id);
uint64_t mbo_stream_size(const mbo_streams* s, int id);
void mbo_streams_clear(mbo_streams* s);
/* appends SEQ_SEPARATOR_MARK to the literal stream: MBGC_Encoder.cpp:489-491 */
void mbo_after_sequence(mbo_streams* s);
/* appends FILE_SEPARATOR_MARK to the flag stream: MBGC_Encoder.cpp:493-496 */
void mbo_after_target(mbo_streams* s, const mbo_pack_params_t* p);

/* MBGC_Encoder.cpp:143-308 (sequential-mode semantics: targetIdx 0, no dissimilar-contig bail-out,
 * refExtLoadedPosArr supplied by the caller for the lazy-decompression pairing rule).
 * matches is modified in place exactly as the reference modifies textMatches (gap-breaking removal,
 * left shifts); *nMatches receives the surviving count.  Returns unmatchedChars (pre-extension). */
uint64_t mbo_process_matches(mbo_streams* out, const mbo_pack_params_t* p, const mbo_matcher* m,
                             mbo_match_t* matches, uint64_t* nMatches, const uint8_t* dest, uint64_t destLen,
                             uint64_t matchingLockPos, const uint64_t* refExtLoadedPosArr, uint64_t nLoadedPos);

/* The same with MBGC_Encoder::processLiteral (mbgccoder/MBGC_Encoder.cpp:128-135, called at :228 and :290 when the contig
 * is NOT proper for a whole-contig reference extension, :211; a developer-build option of the reference, `-N <minimal
 * literal length> -F <context>`): every literal the call appends to the literal stream that is longer than
 * refLiteralMinimalLengthExt gives a range [pos - before, pos + length + after) of the text; targetRefExtensions = the
 * concatenation of those ranges.  UINT64_MAX switches it off (the default, MBGC_Params.h: SIZE_MAX).  *nRanges receives the
 * number of ranges (the first rangeCap are stored). */
typedef struct { uint64_t start, len; } mbo_lit_range_t;
uint64_t mbo_process_matches_lit(mbo_streams* out, const mbo_pack_params_t* p, const mbo_matcher* m,
                                 mbo_match_t* matches, uint64_t* nMatches, const uint8_t* dest, uint64_t destLen,
                                 uint64_t matchingLockPos, const uint64_t* refExtLoadedPosArr, uint64_t nLoadedPos,
                                 uint64_t refLiteralMinimalLengthExt, uint64_t refLiteralBeforeAfterExt,
                                 mbo_lit_range_t* ranges, uint64_t rangeCap, uint64_t* nRanges);

/* coders/ContextAwareMismatchesCoder.cpp:22-27 */
uint8_t mbo_mismatch2code(uint8_t actual, uint8_t mismatch);

/* ---- FASTA ingest: utils/kseq.h:233-274 (kseq_read_lossless_fasta), the reader MBGC uses by default
 * (matching/MultipleGenomeMatchingProcessor.cpp:9-14) -------------------------------------------------------- */

typedef struct {
    uint64_t hdrOff, hdrLen; /* the header line in the input, without '>' and without the line end */
    uint64_t seqOff, seqLen; /* the sequence (line ends removed) in seqOut */
} mbo_fasta_rec_t;

#define MBO_DNA_NOT_WELLFORMED UINT64_MAX /* DNA_NOT_WELLFORMED (-1) in the uint64 member, utils/kseq.h:35,340 */

/* Calls the reader on an in-memory file until it returns a negative value.  Returns the number of records read
 * successfully (recs[0..], sequences concatenated in seqOut, which must hold n bytes); *status = the final negative
 * return value (-1 end of file, -3 not FASTA, -4 DNA not well-formed); *dnaLineLen / *maxLastDnaLineLen = the reader's
 * state after the last call (KSEQ_DNA_LINE_LENGTH reads the former, MultipleGenomeMatchingProcessor.cpp:12-14). */
uint64_t mbo_fasta_parse(const uint8_t* buf, uint64_t n, mbo_fasta_rec_t* recs, uint64_t recCap, uint8_t* seqOut,
                         uint64_t* dnaLineLen, uint64_t* maxLastDnaLineLen, int* status);

/* ---- row a14: SymbolsPackingFacility / VarLenDNACoder (coders/SymbolsPackingFacility.cpp, coders/VarLenDNACoder.cpp) ---- */
int mbo_symbols_per_element(int nSymbols);
uint64_t mbo_symbols_pack(const char* symbols, int nSymbols, const uint8_t* src, uint64_t n, uint8_t* dst);
void mbo_symbols_unpack(const char* symbols, int nSymbols, const uint8_t* packed, uint64_t pos, uint64_t length, uint8_t* dst);
uint64_t mbo_varlen_encode(const char* codes, const uint8_t* src, uint64_t n, uint8_t* dst);
uint64_t mbo_varlen_decode(const char* codes, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dstCap);

/* ---- row f3: SimpleSequenceMatcher::rcMatchSequence over CopMEMMatcher (matching/SimpleSequenceMatcher.cpp:165-176,
 * matching/copmem/CopMEMMatcher.cpp), single-threaded index build (-t 1) ---- */
void mbo_reverse_complement(const uint8_t* src, uint64_t n, uint8_t* dst);
int mbo_copmem_params(uint64_t N, uint32_t L, uint32_t minMatchLength, int* K, int* k1, int* k2, uint32_t* hashSize);
typedef struct { uint64_t src, len, dst; } mbo_rcm_t;   /* TextMatch(posSrcText, length, posDestText) */
/* the list CopMEMMatcher::matchTexts(reverseComplement(seq), destIsSrc, revComplMatching) pushes, in push order; *out is malloc'ed */
int64_t mbo_copmem_rc_matches(const uint8_t* seq, uint64_t N, uint32_t L, uint32_t minMatchLength, mbo_rcm_t** out);
uint64_t mbo_rc_match_sequence(uint8_t* seq, uint64_t N, uint32_t targetMatchLength, uint32_t minMatchLength, uint8_t* mapOff,
                               uint64_t* mapOffLen, uint8_t* mapLen, uint64_t* mapLenLen, uint64_t* nMatchesRaw);

/* ---- row f4: the decoder's match-copy path (mbgccoder/MBGC_Decoder.cpp:243-675), oracle/mbgc_oracle_dec.c ---- */
typedef struct {
    const uint8_t* literals; uint64_t nLiterals;   /* literalStr (after restoreRCMatchedSequence in mode 3) */
    const uint8_t* mapOff;   uint64_t nMapOff;     /* mapOffStream: u32 LE per match that carries its offset */
    const uint8_t* mapOff5;  uint64_t nMapOff5;    /* mapOff5thByteStream (only read when refTotalLength > UINT32_MAX) */
    const uint8_t* mapLen;   uint64_t nMapLen;     /* mapLenStream as archived: u32 each, or frugal u16/u32/u64 words */
    const uint8_t* gapDelta; uint64_t nGapDelta;   /* gapDeltaStr */
    const uint8_t* gapFlags; uint64_t nGapFlags;   /* gapMismatchesFlagsStr */
} mbo_dec_streams_t;
typedef struct mbo_decoder mbo_decoder;
/* the streams are borrowed; refTotalLength as archived (MBGC_Decoder.cpp:1309-1311) */
mbo_decoder* mbo_dec_create(const mbo_pack_params_t* p, uint64_t refTotalLength, int rcInReference,
                            int contigsIndividuallyReversed, const mbo_dec_streams_t* s);
void mbo_dec_destroy(mbo_decoder* d);
/* coders/ContextAwareMismatchesCoder.cpp:29-34 */
uint8_t mbo_code2mismatch(uint8_t actual, uint8_t code);
/* MBGC_Decoder.cpp:243-262 / :264-317 */
void mbo_dec_init_reference(mbo_decoder* d);
uint64_t mbo_dec_decode_reference(mbo_decoder* d, uint32_t seqCount, uint8_t* out, uint64_t* contigLens);
/* MBGC_Decoder.cpp:319-432 (+ extendMatchRight :434-462, extendMatchLeft :464-527) */
uint64_t mbo_dec_decode_sequence(mbo_decoder* d, uint64_t refLockPos, uint8_t* dest, uint64_t* unmatchedOut);
/* MBGC_Decoder.cpp:651-675 */
void mbo_dec_load_ref(mbo_decoder* d, const uint8_t* seqText, uint64_t seqLength, uint64_t refLockPos, int loadRCRef);
/* MBGC_Decoder.cpp:532-649, sequences only */
uint64_t mbo_dec_decode_target(mbo_decoder* d, uint32_t seqCount, int unmatchedFractionFactor, int unmatchedFractionRCFactor,
                               uint64_t refLockPos, uint8_t* out, uint64_t* contigLens, uint64_t* contigUnmatched);
const uint8_t* mbo_dec_ref(const mbo_decoder* d);
uint64_t mbo_dec_ref_pos(const mbo_decoder* d);
int mbo_dec_reached_count(const mbo_decoder* d);
/* literal position, offsets read, 5th bytes read, lengths read, gap deltas read, flag position */
void mbo_dec_cursors(const mbo_decoder* d, uint64_t out6[6]);

#ifdef __cplusplus
}
#endif
#endif
