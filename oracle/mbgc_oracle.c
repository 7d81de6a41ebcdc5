/* oracle/mbgc_oracle.c -- TEST INFRASTRUCTURE ONLY (see mbgc_oracle.h).
 *
 * Single-threaded CPU restatement of MBGC's sparse k-mer exact-match finder and stream packing, written
 * from the behaviour of the reference (file:line cited per function, paths relative to /root/reference).
 * Positions are plain integers instead of the reference's raw pointers; the order of comparisons and the
 * off-by-one conventions of the reference loops are kept literally because they decide the results.
 *
 * Parity status: PINNED against the reference's own build (tests/test_oracle_vs_reference.py,
 * tests/golden/).  Not part of the product; the product path never falls back to it.
 */
#include "mbgc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ */
/* small helpers                                                                                    */

static uint32_t load_u32(const uint8_t* p) {
    uint32_t v;
    memcpy(&v, p, 4);
    return v;
}

/* utils/Hashes.h:28-40.  The accumulator is 64-bit in the reference but only its low 32 bits are
 * returned and neither xor nor multiply moves information downwards, so 32-bit arithmetic is exact. */
uint32_t mbo_hash(const uint8_t* p, int K) {
    uint32_t h = (uint32_t) K;
    for (uint32_t j = 0; j < (uint32_t) K / 4; j++) {
        uint32_t w = load_u32(p + 4 * j) + j;
        h = (h ^ w) * 171717u;
    }
    return h;
}

/* matching/SlidingWindowSparseEMMatcher.cpp:75-87 */
int mbo_kmer_from_L(uint32_t L, uint32_t minMatchLength) {
    int K;
    if (L > 110) K = 56;
    else if (L > 62) K = 44;
    else if (L > 53) K = 40;
    else if (L > 46) K = 36;
    else if (L > 42) K = 32;
    else if (L > 32) K = 28;
    else K = ((int) L / 4 - 1) * 4;
    int KmmL = ((int) minMatchLength / 4 - 1) * 4;
    if (KmmL < K) K = KmmL;
    return K;
}

/* matching/SlidingWindowSparseEMMatcher.cpp:99-103 (HASH_SIZE_MIN_ORDER 24, MAX 31; .h:19-20) */
uint32_t mbo_hash_size(uint64_t maxRefLength, int k1) {
    uint32_t hs;
    int i = 24;
    do {
        hs = ((uint32_t) 1) << (i++);
    } while (i <= 31 && hs < maxRefLength / (uint64_t) k1);
    return hs;
}

/* utils/helper.cpp:312-358: identity except the listed IUPAC letters (either case -> upper-case
 * complement).  The reference fills the identity for CHAR_MIN..CHAR_MAX-1 only, so byte 0x7F maps to 0. */
static uint8_t g_uc_lut[256];
static int g_uc_lut_ready = 0;
static void init_uc_lut(void) {
    if (g_uc_lut_ready) return;
    for (int i = 0; i < 256; i++) g_uc_lut[i] = (uint8_t) i;
    g_uc_lut[0x7F] = 0;
    const char* from = "ACGTNUYRKMBDHVWS";
    const char* to = "TGCANARYMKVHDBSW";
    for (int i = 0; from[i]; i++) {
        g_uc_lut[(uint8_t) from[i]] = (uint8_t) to[i];
        g_uc_lut[(uint8_t) (from[i] + 32)] = (uint8_t) to[i];
    }
    g_uc_lut_ready = 1;
}

/* utils/helper.cpp:405-410 */
void mbo_upper_reverse_complement(const uint8_t* src, uint64_t n, uint8_t* dst) {
    init_uc_lut();
    for (uint64_t i = 0; i < n; i++) dst[n - 1 - i] = g_uc_lut[src[i]];
}

/* matching/SlidingWindowSparseEMMatcher.cpp:146-171.  MULTI2 = 128 samples per OpenMP chunk; the chunked
 * loop and the sequential tail are independent integer ranges (gaps / overlaps are reproduced). */
int mbo_insert_ranges(uint64_t samplingPos, int64_t refPos1, int K, int k1, int64_t* ranges, int cap,
                      uint64_t* newSamplingPos) {
    const int64_t chunk = (int64_t) k1 * 128;
    int n = 0;
    int64_t i1 = (int64_t) samplingPos;
    int64_t chunkedBegin = i1;
    for (; i1 < (refPos1 - K) - chunk; i1 += chunk) {
    }
    if (i1 > chunkedBegin) {
        if (n < cap) {
            ranges[2 * n] = chunkedBegin;
            ranges[2 * n + 1] = i1; /* exclusive; samples chunkedBegin, +k1, ... < i1 */
        }
        n++;
    }
    int64_t t = k1 + ((refPos1 - K - 1) / chunk) * chunk; /* C division: truncates toward zero */
    int64_t tailBegin = t;
    for (; t < refPos1 - K + 1; t += k1) {
    }
    if (t > tailBegin) {
        if (n < cap) {
            ranges[2 * n] = tailBegin;
            ranges[2 * n + 1] = t;
        }
        n++;
    }
    *newSamplingPos = (uint64_t) t;
    return n;
}

/* ------------------------------------------------------------------------------------------------ */
/* matcher state: matching/SlidingWindowSparseEMMatcher.h:28-78                                      */

struct mbo_matcher {
    uint8_t* ref; /* start1 */
    int64_t pos1;
    uint64_t maxRefLength;
    int reachedRefLengthCount;
    int L, K, k1, k2, skipMargin;
    int k1ord; /* -1: plain variant (position stored as is) */
    uint32_t hashSize, hashMask;
    int LK2, K_PLUS_LK24;
    uint64_t swSize, swEnd;
    int circularBuffer;
    uint64_t* locks; /* workersSwEndPositions (deque) */
    size_t nLocks, capLocks;
    uint32_t* ht;
    uint64_t samplingPos;
    mbo_match_t* res;
    size_t nRes, capRes;
    uint64_t stats[4];
};

static const uint64_t REF_SHIFT = 1; /* .h:14 */

static uint64_t ht_encode(const mbo_matcher* m, uint64_t pos) { return m->k1ord >= 0 ? pos >> m->k1ord : pos; }
static uint64_t ht_decode(const mbo_matcher* m, uint64_t code) { return m->k1ord >= 0 ? code << m->k1ord : code; }

mbo_matcher* mbo_create(uint64_t refLengthLimit, uint32_t L, int k1, int k2, int skipMargin) {
    init_uc_lut();
    mbo_matcher* m = (mbo_matcher*) calloc(1, sizeof(mbo_matcher));
    /* the reference additionally caps the allocation at 60 % of RAM (.cpp:331-334); not restated */
    m->maxRefLength = refLengthLimit;
    m->ref = (uint8_t*) calloc(refLengthLimit + 64, 1);
    m->ref[0] = 0;
    m->L = (int) L;
    m->skipMargin = skipMargin;
    m->pos1 = (int64_t) REF_SHIFT;
    m->swEnd = m->maxRefLength;
    m->circularBuffer = 1;
    m->swSize = m->maxRefLength / 16; /* SW_WIDTH_FACTOR, .h:47, .cpp:339 */
    /* .cpp:342-344: minMatchLength defaults to UINT32_MAX and is clamped to L */
    m->K = mbo_kmer_from_L(L, L);
    m->k1 = k1;
    m->k2 = k2;
    m->LK2 = (m->L - m->K) / 2;
    m->K_PLUS_LK24 = m->K + m->LK2 - 4;
    m->hashSize = mbo_hash_size(m->maxRefLength, k1);
    m->hashMask = m->hashSize - 1;
    /* MultipleGenomeMatchingProcessor.cpp:170-176: ctz(k1) != 0 selects the Exp variant (.cpp:494-519) */
    if (__builtin_ctz((uint32_t) k1)) {
        m->k1ord = __builtin_ctz((uint32_t) k1);
        m->samplingPos = (uint64_t) k1; /* .cpp:503 */
    } else {
        m->k1ord = -1;
        m->samplingPos = REF_SHIFT; /* .h:78 */
    }
    m->ht = (uint32_t*) calloc(m->hashSize, sizeof(uint32_t));
    return m;
}

void mbo_destroy(mbo_matcher* m) {
    if (!m) return;
    free(m->ref);
    free(m->ht);
    free(m->locks);
    free(m->res);
    free(m);
}

/* .h:93-97 */
void mbo_disable_sliding_window(mbo_matcher* m) {
    m->swSize = 0;
    m->swEnd = m->circularBuffer ? 0 : m->maxRefLength;
}
void mbo_disable_circular_buffer(mbo_matcher* m) {
    m->circularBuffer = 0;
    m->swEnd = m->maxRefLength;
}
void mbo_set_sliding_window_size(mbo_matcher* m, int factor) { m->swSize = m->maxRefLength / (uint8_t) factor; }

/* .cpp:146-171, single thread: chunked part in ascending order, then the tail; later store wins */
static void index_pass(mbo_matcher* m) {
    int64_t ranges[4];
    uint64_t newSampling;
    int n = mbo_insert_ranges(m->samplingPos, m->pos1, m->K, m->k1, ranges, 2, &newSampling);
    for (int r = 0; r < n; r++)
        for (int64_t i = ranges[2 * r]; i < ranges[2 * r + 1]; i += m->k1)
            m->ht[mbo_hash(m->ref + i, m->K) & m->hashMask] = (uint32_t) ht_encode(m, (uint64_t) i);
    m->samplingPos = newSampling;
}

/* .cpp:402-437 (the reference recurses on the remainder; this is the same thing as a loop) */
static void load_segment(mbo_matcher* m, const uint8_t* text, uint64_t len, int rc, int addSep, uint8_t sep) {
    while (len != 0) {
        if ((uint64_t) m->pos1 == m->maxRefLength && m->swEnd != m->maxRefLength) {
            m->reachedRefLengthCount++;
            m->pos1 = (int64_t) REF_SHIFT;
            m->samplingPos = REF_SHIFT;
        }
        uint64_t tmpEnd = m->swEnd;
        uint64_t tmpLen = len;
        uint64_t tmpMax = tmpEnd < (uint64_t) m->pos1 ? m->maxRefLength : tmpEnd;
        if ((uint64_t) m->pos1 + tmpLen > tmpMax) tmpLen = tmpMax - (uint64_t) m->pos1;
        if (rc)
            mbo_upper_reverse_complement(text + len - tmpLen, tmpLen, m->ref + m->pos1);
        else
            memcpy(m->ref + m->pos1, text, tmpLen);
        if (addSep && (uint64_t) m->pos1 + tmpLen == m->swEnd) m->ref[m->swEnd - 1] = sep;
        m->pos1 += (int64_t) tmpLen;
        index_pass(m);
        if (!rc) text += tmpLen;
        len = ((uint64_t) m->pos1 == tmpEnd) ? 0 : len - tmpLen;
    }
}

/* .cpp:453-458 */
void mbo_load_ref(mbo_matcher* m, const uint8_t* text, uint64_t len, int loadRC, int addSep, uint8_t sep) {
    load_segment(m, text, len, 0, addSep, sep);
    if (loadRC) load_segment(m, text, len, 1, addSep, sep);
}

/* .cpp:439-451 */
void mbo_load_separator(mbo_matcher* m, uint8_t sep) {
    if ((uint64_t) m->pos1 == m->maxRefLength && m->swEnd != m->maxRefLength) {
        m->reachedRefLengthCount++;
        m->pos1 = (int64_t) REF_SHIFT;
        m->samplingPos = REF_SHIFT;
    }
    if ((uint64_t) m->pos1 == m->maxRefLength) return;
    if ((uint64_t) m->pos1 == m->swEnd)
        m->ref[m->pos1 - 1] = sep;
    else
        m->ref[m->pos1++] = sep;
}

/* .cpp:361-378 */
uint64_t mbo_acquire_lock(mbo_matcher* m) {
    if (m->swSize == 0 || !m->circularBuffer) return m->swEnd;
    uint64_t w = (uint64_t) m->pos1 + m->swSize;
    if (m->reachedRefLengthCount || w > m->maxRefLength) {
        if (w > m->maxRefLength) w -= m->maxRefLength - REF_SHIFT;
    } else
        w = m->maxRefLength;
    if (m->nLocks == 0) m->swEnd = w;
    if (m->nLocks == m->capLocks) {
        m->capLocks = m->capLocks ? 2 * m->capLocks : 16;
        m->locks = (uint64_t*) realloc(m->locks, m->capLocks * sizeof(uint64_t));
    }
    m->locks[m->nLocks++] = w;
    return w;
}

/* .cpp:380-400; returns 0 on success, -1 on the reference's "Invalid worker lock value" exit */
int mbo_release_lock(mbo_matcher* m, uint64_t lockValue) {
    if (m->swSize == 0 || !m->circularBuffer) return 0;
    size_t i = 0;
    while (i < m->nLocks && m->locks[i] != lockValue) i++;
    if (i == m->nLocks) return -1;
    if (i == 0) {
        size_t drop = 0;
        do {
            drop++;
        } while (drop < m->nLocks && m->locks[drop] == MBO_NO_LOCK);
        memmove(m->locks, m->locks + drop, (m->nLocks - drop) * sizeof(uint64_t));
        m->nLocks -= drop;
        m->swEnd = m->nLocks == 0 ? m->swEnd : m->locks[0];
    } else
        m->locks[i] = MBO_NO_LOCK;
    return 0;
}

/* .h:104-113 */
const uint8_t* mbo_ref(const mbo_matcher* m) { return m->ref; }
uint64_t mbo_max_ref_length(const mbo_matcher* m) { return m->maxRefLength; }
uint64_t mbo_ref_length(const mbo_matcher* m) { return m->reachedRefLengthCount ? m->maxRefLength : (uint64_t) m->pos1; }
uint64_t mbo_loading_position(const mbo_matcher* m) { return (uint64_t) m->pos1; }
uint64_t mbo_loaded_ref_length(const mbo_matcher* m) {
    return (uint64_t) m->reachedRefLengthCount * (m->maxRefLength - REF_SHIFT) + ((uint64_t) m->pos1 - REF_SHIFT);
}
void mbo_set_position(mbo_matcher* m, uint64_t refPos, int reachedCount) {
    m->pos1 = (int64_t) refPos;
    m->reachedRefLengthCount = reachedCount;
}
const uint32_t* mbo_table(const mbo_matcher* m) { return m->ht; }
uint64_t mbo_table_size(const mbo_matcher* m) { return m->hashSize; }
uint64_t mbo_sampling_pos(const mbo_matcher* m) { return m->samplingPos; }
int mbo_kmer(const mbo_matcher* m) { return m->K; }
const mbo_match_t* mbo_matches(const mbo_matcher* m) { return m->res; }
void mbo_last_stats(const mbo_matcher* m, uint64_t out4[4]) { memcpy(out4, m->stats, sizeof(m->stats)); }

static void res_push(mbo_matcher* m, uint64_t src, uint64_t len, uint64_t dst) {
    if (m->nRes == m->capRes) {
        m->capRes = m->capRes ? 2 * m->capRes : 1024;
        m->res = (mbo_match_t*) realloc(m->res, m->capRes * sizeof(mbo_match_t));
    }
    m->res[m->nRes].src = src;
    m->res[m->nRes].len = len;
    m->res[m->nRes].dst = dst;
    m->nRes++;
}

/* .cpp:181-321, checkOverlaps = true, PAIRED_MATCH_LENGTH_LOSS_LIMIT = 0 (its branches are dead code).
 * T = target text, R = reference buffer.  p1/p2 play the role of the reference's pointers of the same
 * name as integer positions in R / T; p2 may become -1 exactly where the reference points at start2-1. */
uint64_t mbo_match(mbo_matcher* m, const uint8_t* T, uint64_t N2, uint32_t minMatchLength, uint64_t lockPos,
                   int usePrefilter) {
    if (minMatchLength < (uint32_t) m->K) return UINT64_MAX; /* .cpp:480-483 */
    const uint8_t* R = m->ref;
    const int K = m->K, k2 = m->k2;
    m->nRes = 0;
    memset(m->stats, 0, sizeof(m->stats));
    uint32_t l1 = 0, l2 = 0, r1 = 0, r2 = 0;
    const int64_t OVERLAP_MATCH_MAX_LENGTH = 1 << 13; /* .h:18 */

    for (int64_t i2 = 0; i2 + K < (int64_t) N2 + 1; i2 += k2) {
        uint32_t slot = mbo_hash(T + i2, K) & m->hashMask;
        if (i2 - m->LK2 >= 0) l2 = load_u32(T + i2 - m->LK2);
        if ((uint64_t) (i2 + m->K_PLUS_LK24 + 4) <= N2) r2 = load_u32(T + i2 + m->K_PLUS_LK24);
        m->stats[0]++;
        if (m->ht[slot] == 0) continue;
        m->stats[1]++;

        const int64_t s = (int64_t) ht_decode(m, m->ht[slot]); /* curr1 */
        const int64_t swStart = m->pos1;
        const int64_t swStop = lockPos != MBO_NO_LOCK ? (int64_t) lockPos : m->pos1;
        const int endsBeforeSwStart = s + K < swStart;
        const int startsBeforeSwStop = s < swStop;
        if (swStart <= swStop) {
            if (!endsBeforeSwStart && startsBeforeSwStop) continue;
        } else if (!endsBeforeSwStart || startsBeforeSwStop)
            continue;
        const int64_t hi = endsBeforeSwStart ? swStart : (int64_t) mbo_ref_length(m); /* tmpEnd1 */
        const int64_t lo = startsBeforeSwStop ? 0 : swStop;                           /* tmpStart1 */
        if (usePrefilter) {
            /* s >= k1 > LK2 for the Exp variant; the plain variant can read before the buffer in the
             * reference (garbage that cannot match): treated as "no match" here. */
            l1 = s - m->LK2 >= 0 ? load_u32(R + s - m->LK2) : ~l2;
            r1 = load_u32(R + s + m->K_PLUS_LK24);
            if (!(r1 == r2 || l1 == l2)) continue;
        }

        /* right end: .cpp:227-246 (the 64-byte block compares are an unrolled form of this loop) */
        int64_t p1 = s + K - 1, p2 = i2 + K - 1;
        while (++p1 != hi && ++p2 != (int64_t) N2 && R[p1] == T[p2]) {
        }
        const int64_t right1 = p1;
        p1 = s;
        p2 = i2;

        /* left start with overlap resolution against the accepted matches: .cpp:251-291 */
        int64_t keep = (int64_t) m->nRes; /* resSizeWithoutOverlapped */
        while (keep--) {
            const mbo_match_t* last = &m->res[keep];
            const int64_t lastEnd = (int64_t) (last->dst + last->len);
            if (lastEnd < p2) {
                while (p1 != lo && p2 > lastEnd - 1 && R[p1] == T[p2]) {
                    p1--;
                    p2--;
                }
                if (p2 > lastEnd - 1) break;
                p1++;
                p2++;
            }
            const int64_t lastDelta = p2 - (int64_t) last->dst;
            if (p1 - lo < lastDelta || (int64_t) last->len > OVERLAP_MATCH_MAX_LENGTH ||
                memcmp(T + p2 - lastDelta, R + p1 - lastDelta, (size_t) lastDelta) != 0) {
                p1--;
                p2--;
                break;
            }
            p2 -= lastDelta;
            p1 -= lastDelta;
        }
        if (keep < 0) {
            while (p1 != lo && p2 > -1 && R[p1] == T[p2]) {
                p1--;
                p2--;
            }
        } else {
            const int64_t overlap = (int64_t) (m->res[keep].dst + m->res[keep].len) - (p2 + 1);
            if (overlap > 0) {
                p1 += overlap;
                p2 += overlap;
            }
        }
        ++keep;

        /* accept: .cpp:298-313 */
        if (right1 - p1 > (int64_t) minMatchLength && memcmp(R + s, T + i2, (size_t) K) == 0) {
            m->stats[2]++;
            m->stats[3]++;
            m->nRes = (size_t) keep;
            res_push(m, (uint64_t) (p1 + 1), (uint64_t) (right1 - p1 - 1), (uint64_t) (p2 + 1));
            int skip = (int) (((p2 + right1 - p1) - i2) / k2 * k2);
            skip -= skip > m->skipMargin ? m->skipMargin : skip;
            if (skip) i2 += skip - k2;
        } else if (memcmp(R + s, T + i2, (size_t) K) == 0)
            m->stats[2]++;
    }
    return m->nRes;
}

/* ------------------------------------------------------------------------------------------------ */
/* stream packing: mbgccoder/MBGC_Encoder.cpp:143-427                                               */

typedef struct {
    uint8_t* p;
    uint64_t n, cap;
} bytes_t;

struct mbo_streams {
    bytes_t s[MBO_S_COUNT];
};

static void bput(bytes_t* b, const void* src, uint64_t n) {
    if (b->n + n > b->cap) {
        uint64_t c = b->cap ? b->cap : 4096;
        while (c < b->n + n) c *= 2;
        b->p = (uint8_t*) realloc(b->p, c);
        b->cap = c;
    }
    if (n) memcpy(b->p + b->n, src, n);
    b->n += n;
}
static void bput1(bytes_t* b, uint8_t v) { bput(b, &v, 1); }

mbo_streams* mbo_streams_create(void) { return (mbo_streams*) calloc(1, sizeof(mbo_streams)); }
void mbo_streams_destroy(mbo_streams* s) {
    if (!s) return;
    for (int i = 0; i < MBO_S_COUNT; i++) free(s->s[i].p);
    free(s);
}
const uint8_t* mbo_stream_data(const mbo_streams* s, int id) { return s->s[id].p; }
uint64_t mbo_stream_size(const mbo_streams* s, int id) { return s->s[id].n; }
void mbo_streams_clear(mbo_streams* s) {
    for (int i = 0; i < MBO_S_COUNT; i++) s->s[i].n = 0;
}

/* mbgccoder/MBGC_Params.h:45-48 */
#define MATCH_MARK ((uint8_t) ('%' + 128))
#define SEQ_SEPARATOR_MARK ((uint8_t) ('"' + 128))
#define FILE_SEPARATOR_MARK ((uint8_t) (';' + 128))
#define REF_REGION_SEPARATOR 0 /* matching/MGMP_Params.h:14 */
#define MAX_GAP_DEPTH 128      /* mbgccoder/MBGC_Params.h:50 */
#define MAX_EXTEND_MATCH_LEFT_LENGTH (1 << 24) /* mbgccoder/MBGC_Params.h:55 */

void mbo_after_sequence(mbo_streams* s) { bput1(&s->s[MBO_S_LITERALS], SEQ_SEPARATOR_MARK); }
void mbo_after_target(mbo_streams* s, const mbo_pack_params_t* p) {
    if (p->enableExtensionsWithMismatches) bput1(&s->s[MBO_S_GAPFLAGS], FILE_SEPARATOR_MARK);
}

/* coders/ContextAwareMismatchesCoder.h:13-17, .cpp:22-27.  sym2val: A,C,G,T,N -> 0..4, everything else
 * > 4 (the reference indexes with a signed char; bytes >= 0x80 are out of its domain). */
uint8_t mbo_mismatch2code(uint8_t actual, uint8_t mismatch) {
    static const uint8_t mis2code[5][5] = {{'A', 2, 0, 1, 3}, {1, 'C', 2, 0, 3}, {0, 2, 'G', 1, 3},
                                           {1, 0, 2, 'T', 3}, {1, 2, 3, 0, 'N'}};
    int a, b;
    switch (actual) { case 'A': a = 0; break; case 'C': a = 1; break; case 'G': a = 2; break;
                      case 'T': a = 3; break; case 'N': a = 4; break; default: a = 255; }
    switch (mismatch) { case 'A': b = 0; break; case 'C': b = 1; break; case 'G': b = 2; break;
                        case 'T': b = 3; break; case 'N': b = 4; break; default: b = 255; }
    if (actual == mismatch || a > 4 || b > 4) return mismatch;
    return mis2code[a][b];
}

/* utils/helper.cpp:237-246 */
static void write_u64_frugal(bytes_t* b, uint64_t value) {
    uint16_t w = value < UINT16_MAX ? (uint16_t) value : UINT16_MAX;
    bput(b, &w, 2);
    if (value >= UINT16_MAX) {
        uint32_t y = value < UINT32_MAX ? (uint32_t) value : UINT32_MAX;
        bput(b, &y, 4);
        if (value >= UINT32_MAX) bput(b, &value, 8);
    }
}

/* matching/TextMatchers.h:42-50 */
static int paired(const mbo_match_t* a, const mbo_match_t* b) { return a->src + b->dst == b->src + a->dst; }
static int paired_break(const mbo_match_t* a, const mbo_match_t* b, uint64_t brk) {
    return a->src + b->dst == b->src + a->dst &&
           ((a->src > brk && b->src > brk) || (a->src < brk && b->src < brk));
}

/* MBGC_Encoder.cpp:137-141 */
static uint64_t match_loaded_pos(const mbo_matcher* m, uint64_t pos, const uint64_t* arr, uint64_t n) {
    while (pos + (mbo_ref_length(m) - REF_SHIFT) < arr[n - 1]) pos += (mbo_ref_length(m) - REF_SHIFT);
    return pos;
}

/* MBGC_Encoder.cpp:373-427 */
static uint64_t extend_left(mbo_streams* out, const mbo_pack_params_t* p, const mbo_matcher* m, const uint8_t* dest,
                            uint64_t length, const mbo_match_t* match, uint64_t lockPos) {
    const uint8_t* R = m->ref;
    const int64_t srcMatch = (int64_t) match->src;
    int64_t srcGuard = 1;
    if (srcGuard < srcMatch - MAX_EXTEND_MATCH_LEFT_LENGTH) srcGuard = srcMatch - MAX_EXTEND_MATCH_LEFT_LENGTH;
    /* srcLock = ref + lockPos is a wrapped pointer when lockPos == SIZE_MAX (ref - 1): compares below every
     * valid position, i.e. as -1 */
    const int64_t srcLock = (int64_t) lockPos;
    if (srcGuard < srcLock && srcLock <= srcMatch) srcGuard = srcLock;
    int guardKnown = 1;
    if (srcGuard < srcMatch - (int64_t) length) {
        srcGuard = srcMatch - (int64_t) length;
        guardKnown = 0;
    }
    if (srcGuard == srcMatch) return 0;
    int64_t srcPtr = srcMatch - 1;
    int64_t gapPtr = (int64_t) match->dst - 1;
    int validSrc = !p->lazyDecompressionSupport || R[srcPtr] != REF_REGION_SEPARATOR;
    bput1(&out->s[MBO_S_LITERALS], p->mismatchesWithExclusion && validSrc
                                       ? mbo_mismatch2code(R[srcPtr], dest[gapPtr]) : dest[gapPtr]);
    int score = p->mmsMismatchesInitialScore;
    while (validSrc && srcPtr > srcGuard && score < p->mmsMismatchesScoreThreshold) {
        --gapPtr;
        --srcPtr;
        int mismatch = dest[gapPtr] != R[srcPtr];
        if (p->lazyDecompressionSupport && R[srcPtr] == REF_REGION_SEPARATOR) {
            validSrc = 0;
            srcPtr++;
            gapPtr++;
            break;
        }
        bput1(&out->s[MBO_S_GAPFLAGS], mismatch ? 1 : 0);
        if (mismatch) {
            score += p->mmsMismatchPenalty;
            bput1(&out->s[MBO_S_LITERALS], p->mismatchesWithExclusion
                                               ? mbo_mismatch2code(R[srcPtr], dest[gapPtr]) : dest[gapPtr]);
        } else {
            score -= p->mmsMatchBonus;
            if (score < 0) score = 0;
        }
    }
    while (!validSrc && srcPtr > srcGuard && score < p->mmsMismatchesScoreThreshold) {
        srcPtr--;
        bput1(&out->s[MBO_S_GAPFLAGS], 1);
        bput1(&out->s[MBO_S_LITERALS], dest[--gapPtr]);
        score += p->mmsMismatchPenalty;
    }
    if ((srcPtr != srcGuard || !guardKnown) && score < p->mmsMismatchesScoreThreshold)
        bput1(&out->s[MBO_S_GAPFLAGS], 1);
    return (uint64_t) (srcMatch - srcPtr);
}

/* MBGC_Encoder.cpp:310-371.  gap = destination bytes right after `match` (length of them in *length). */
static uint64_t extend_right(mbo_streams* out, const mbo_pack_params_t* p, const mbo_matcher* m, const uint8_t* gap,
                             const mbo_match_t* core, const mbo_match_t* match, uint64_t length, int isGap,
                             int gapStart, int gapMiddle, int gapEnd) {
    if (length == 0) {
        if (gapMiddle) bput1(&out->s[MBO_S_GAPFLAGS], 1);
        return 0;
    }
    const uint8_t* R = m->ref;
    int64_t gapPos = 0;
    int64_t srcPtr = (int64_t) (isGap ? core->src + (match->dst + match->len) - core->dst : match->src + match->len);
    const int64_t loadingPos = (int64_t) mbo_loading_position(m);
    const int64_t srcGuard = srcPtr + (int64_t) length;
    int64_t validGuard = srcPtr + (int64_t) length;
    if (srcPtr == loadingPos) validGuard = srcPtr;
    if (!isGap) {
        const int64_t srcEnd = (int64_t) m->maxRefLength;
        if (validGuard > srcEnd) validGuard = srcEnd;
        if (srcPtr <= loadingPos && loadingPos < validGuard) validGuard = loadingPos;
    }
    if (gapStart || !isGap) {
        if (p->lazyDecompressionSupport && R[srcPtr] == REF_REGION_SEPARATOR) validGuard = srcPtr;
        bput1(&out->s[MBO_S_LITERALS], p->mismatchesWithExclusion && srcPtr < validGuard
                                           ? mbo_mismatch2code(R[srcPtr], gap[gapPos]) : gap[gapPos]);
        gapPos++;
    } else
        srcPtr--;
    int score = p->mmsMismatchesInitialScore;
    while (++srcPtr < validGuard && (!p->lazyDecompressionSupport || R[srcPtr] != REF_REGION_SEPARATOR) &&
           (isGap || score < p->mmsMismatchesScoreThreshold)) {
        int mismatch = gap[gapPos] != R[srcPtr];
        bput1(&out->s[MBO_S_GAPFLAGS], mismatch ? 1 : 0);
        if (mismatch) {
            score += p->mmsMismatchPenalty;
            bput1(&out->s[MBO_S_LITERALS], p->mismatchesWithExclusion
                                               ? mbo_mismatch2code(R[srcPtr], gap[gapPos]) : gap[gapPos]);
        } else {
            score -= p->mmsMatchBonus;
            if (score < 0) score = 0;
        }
        gapPos++;
    }
    while (srcPtr++ < srcGuard && (isGap || score < p->mmsMismatchesScoreThreshold)) {
        bput1(&out->s[MBO_S_GAPFLAGS], 1);
        bput1(&out->s[MBO_S_LITERALS], gap[gapPos++]);
        score += p->mmsMismatchPenalty;
    }
    if ((isGap && !gapEnd) || (!isGap && score < p->mmsMismatchesScoreThreshold)) bput1(&out->s[MBO_S_GAPFLAGS], 1);
    return (uint64_t) gapPos;
}

/* MBGC_Encoder.cpp:143-308 */
/* MBGC_Encoder::processLiteral (mbgccoder/MBGC_Encoder.cpp:128-135): a literal longer than refLiteralMinimalLengthExt,
 * with refLiteralBeforeAfterExt bytes of context on either side, is appended to targetRefExtensions -- recorded here as a
 * range of the text instead of copied bytes. */
static void process_literal(uint32_t pos, uint64_t length, uint64_t destLen, uint64_t minLen, uint64_t beforeAfter,
                            mbo_lit_range_t* ranges, uint64_t cap, uint64_t* nRanges) {
    if (length > minLen) {                                       /* MGMP_Params.h:175-177 */
        const int before = beforeAfter < pos ? (int) beforeAfter : (int) pos;
        const int after = (pos + length + beforeAfter < destLen) ? (int) beforeAfter : (int) (destLen - pos - length);
        if (*nRanges < cap) {
            ranges[*nRanges].start = (uint64_t) pos - (uint64_t) before;
            ranges[*nRanges].len = length + (uint64_t) before + (uint64_t) after;
        }
        (*nRanges)++;
    }
}

uint64_t mbo_process_matches(mbo_streams* out, const mbo_pack_params_t* p, const mbo_matcher* m, mbo_match_t* tm,
                             uint64_t* nMatches, const uint8_t* dest, uint64_t destLen, uint64_t lockPos,
                             const uint64_t* loadedArr, uint64_t nLoaded) {
    uint64_t nRanges = 0;
    return mbo_process_matches_lit(out, p, m, tm, nMatches, dest, destLen, lockPos, loadedArr, nLoaded, UINT64_MAX, 0, NULL, 0,
                                   &nRanges);
}

uint64_t mbo_process_matches_lit(mbo_streams* out, const mbo_pack_params_t* p, const mbo_matcher* m, mbo_match_t* tm,
                                 uint64_t* nMatches, const uint8_t* dest, uint64_t destLen, uint64_t lockPos,
                                 const uint64_t* loadedArr, uint64_t nLoaded, uint64_t refLiteralMinimalLengthExt,
                                 uint64_t refLiteralBeforeAfterExt, mbo_lit_range_t* ranges, uint64_t rangeCap,
                                 uint64_t* nRanges) {
    const uint8_t* R = m->ref;
    const int64_t n0 = (int64_t) *nMatches;
    uint32_t pos = 0;
    int64_t unmatched = 0;
    int64_t jj = 0, j;
    /* pass 1: drop gap-breaking matches (:153-199) */
    for (j = 0; j < n0; j++, jj++) {
        mbo_match_t* match = &tm[j];
        if (p->enableExtensionsWithMismatches && j + 1 < n0) {
            if (jj > 0 && paired(&tm[j + 1], &tm[jj - 1]) && !paired(match, &tm[jj - 1]) &&
                match->len < (uint64_t) p->gapBreakingMatchMinLength) {
                int64_t leftExt = 0;
                if (match->dst + match->len == tm[j + 1].dst) {
                    int64_t sp = (int64_t) tm[j + 1].src;
                    int64_t dp = (int64_t) tm[j + 1].dst;
                    while (dest[--dp] == R[--sp]) leftExt++;
                    tm[j + 1].src -= (uint64_t) leftExt;
                    tm[j + 1].dst -= (uint64_t) leftExt;
                    tm[j + 1].len += (uint64_t) leftExt;
                }
                jj--;
                continue;
            }
        }
        unmatched += (int64_t) (match->dst - pos);
        pos = (uint32_t) (match->dst + match->len);
        tm[jj] = tm[j];
    }
    const int64_t n = jj;
    *nMatches = (uint64_t) n;
    unmatched += (int64_t) (destLen - pos);
    /* :211 loadLiteralsToRef = !isContigProperForRefExtension(destLen, unmatchedChars, factor)  (MGMP_Params.h:179-187) */
    const int loadLiteralsToRef = refLiteralMinimalLengthExt != UINT64_MAX &&
                                  !((uint64_t) unmatched * (uint64_t) p->unmatchedFractionFactor > destLen);
    *nRanges = 0;

    /* next region loading positions live beside the matches (TextMatch::nextSrcRegionLoadingPos) */
    uint64_t* nextLoad = (uint64_t*) calloc((size_t) (n > 0 ? n : 1), sizeof(uint64_t));
    uint8_t pairedGap[MAX_GAP_DEPTH];
    memset(pairedGap, 0, sizeof(pairedGap));
    const int64_t NO_GAP = -1;
    int64_t gapStartIdx = NO_GAP, gapEndIdx = NO_GAP;
    int gapCurIdx = 0;
    pos = 0;
    int isGap = 0;
    bytes_t* lit = &out->s[MBO_S_LITERALS];
    for (j = 0; j < n; j++) {
        mbo_match_t* match = &tm[j];
        uint64_t literalsLeft = match->dst - pos;
        if (p->enableExtensionsWithMismatches) {
            if (!isGap && literalsLeft) literalsLeft -= extend_left(out, p, m, dest, literalsLeft, match, lockPos);
            if (j == gapEndIdx) {
                gapStartIdx = NO_GAP;
                gapEndIdx = NO_GAP;
            }
        }
        bput(lit, dest + pos, literalsLeft);
        if (loadLiteralsToRef)                                    /* :228 */
            process_literal(pos, literalsLeft, destLen, refLiteralMinimalLengthExt, refLiteralBeforeAfterExt, ranges, rangeCap, nRanges);
        if (!pairedGap[gapCurIdx]) {
            uint32_t off = (uint32_t) match->src;
            bput(&out->s[MBO_S_MAPOFF], &off, 4);
            if (p->enable40bitReference) bput1(&out->s[MBO_S_MAPOFF5], (uint8_t) (match->src >> 32));
        }
        if (p->frugal64bitLenEncoding)
            write_u64_frugal(&out->s[MBO_S_MAPLEN], match->len);
        else {
            uint32_t l32 = (uint32_t) match->len;
            bput(&out->s[MBO_S_MAPLEN], &l32, 4);
        }
        bput1(lit, MATCH_MARK);
        pos = (uint32_t) (match->dst + match->len);

        literalsLeft = (j + 1 < n ? tm[j + 1].dst : destLen) - pos;
        int gapIdx = (gapCurIdx + 1) % MAX_GAP_DEPTH;
        int gCnt = (int) (n - j - 1) < p->gapDepthOffsetEncoding ? (int) (n - j - 1) : p->gapDepthOffsetEncoding;
        int reduce = 0, g;
        for (g = 1; g <= gCnt; g++, gapIdx = (gapIdx + 1) % MAX_GAP_DEPTH) {
            if (pairedGap[gapIdx] ||
                (!p->lazyDecompressionSupport && g == 1 && gapStartIdx == NO_GAP && literalsLeft == 0)) {
                reduce++;
                continue;
            }
            if (paired_break(match, &tm[j + g], lockPos)) {
                if (p->lazyDecompressionSupport) {
                    if (!nextLoad[j]) {
                        /* std::upper_bound(arr, arr+n, loadedPos(match.src)) */
                        uint64_t key = match_loaded_pos(m, match->src, loadedArr, nLoaded);
                        uint64_t lo = 0, hi = nLoaded;
                        while (lo < hi) {
                            uint64_t mid = (lo + hi) / 2;
                            if (loadedArr[mid] <= key) lo = mid + 1; else hi = mid;
                        }
                        nextLoad[j] = lo == nLoaded ? UINT64_MAX : loadedArr[lo];
                    }
                    if (match_loaded_pos(m, tm[j + g].src, loadedArr, nLoaded) >= nextLoad[j]) continue;
                    nextLoad[j + g] = nextLoad[j];
                }
                pairedGap[gapIdx] = 1;
                if (p->enableExtensionsWithMismatches && gapEndIdx <= j + g && g <= p->gapDepthMismatchesEncoding) {
                    gapStartIdx = j;
                    gapEndIdx = j + g;
                }
                break;
            }
        }
        if (gCnt) bput1(&out->s[MBO_S_GAPDELTA], (uint8_t) (g <= gCnt ? g - reduce : 0));
        pairedGap[gapCurIdx] = 0;
        gapCurIdx = (gapCurIdx + 1) % MAX_GAP_DEPTH;

        const int gapStart = gapStartIdx == j;
        const int gapEnd = gapEndIdx == j + 1;
        const int gapMiddle = gapStartIdx < j && j + 1 < gapEndIdx;
        isGap = gapStart || gapMiddle || gapEnd;
        if (p->enableExtensionsWithMismatches) {
            const mbo_match_t* core = isGap ? &tm[gapStartIdx] : match;
            uint64_t ext = extend_right(out, p, m, dest + pos, core, match, literalsLeft, isGap, gapStart, gapMiddle,
                                        gapEnd);
            pos += (uint32_t) ext;
        }
    }
    bput(lit, dest + pos, destLen - pos);
    if (loadLiteralsToRef)                                        /* :290 */
        process_literal(pos, destLen - pos, destLen, refLiteralMinimalLengthExt, refLiteralBeforeAfterExt, ranges, rangeCap, nRanges);
    free(nextLoad);
    return (uint64_t) unmatched;
}

/* ---- FASTA ingest ------------------------------------------------------------------------------------------- */

/* utils/kseq.h:233-274.  The stream functions (ks_getc :61-72, ks_getuntil2 :100-149 with KS_SEP_LINE and
 * loosy = false: a line ends at '\n' only, '\r' is data) are restated on a cursor into the buffer. */
uint64_t mbo_fasta_parse(const uint8_t* buf, uint64_t n, mbo_fasta_rec_t* recs, uint64_t recCap, uint8_t* seqOut,
                         uint64_t* dnaLineLenOut, uint64_t* maxLastOut, int* status) {
    uint64_t pos = 0, nrec = 0, outl = 0;
    uint64_t dna = 0 /* DNA_LINE_LENGTH_UNKNOWN */, maxLast = 0;
    int last_char = 0;
    *status = -1;
    for (;;) {
        if (last_char == 0) { /* :238-242 */
            if (pos >= n) { *status = -1; break; }
            int c = buf[pos++];
            if (c != '>') { *status = -3; break; }
            last_char = c;
        }
        /* :244 ks_getuntil(KS_SEP_LINE, name): nothing left -> -1 */
        if (pos >= n) { *status = -1; break; }
        const uint64_t hdrOff = pos;
        const uint8_t* e = (const uint8_t*) memchr(buf + pos, '\n', n - pos);
        const uint64_t hdrEnd = e ? (uint64_t) (e - buf) : n;
        pos = e ? hdrEnd + 1 : n;
        uint64_t seql = 0, b = 0;
        int c = -1;
        while (pos < n) { /* :250-259; ks_getc at the end of the buffer returns -1 and ends the loop */
            c = buf[pos++];
            if (c == '>') break;
            if (c == '\n') { dna = MBO_DNA_NOT_WELLFORMED; c = -1; continue; }
            if (seql > b) {
                const uint64_t d = seql - b;
                dna = dna == 0 ? d : (dna == d ? d : MBO_DNA_NOT_WELLFORMED);
            }
            b = seql;
            seqOut[outl + seql++] = (uint8_t) c;
            const uint8_t* le = (const uint8_t*) memchr(buf + pos, '\n', n - pos);
            const uint64_t lineEnd = le ? (uint64_t) (le - buf) : n;
            memcpy(seqOut + outl + seql, buf + pos, lineEnd - pos);
            seql += lineEnd - pos;
            pos = le ? lineEnd + 1 : n;
            c = -1;
        }
        if (seql > b) { /* :260-263 */
            const uint64_t d = seql - b;
            if (d > maxLast) maxLast = d;
        }
        dna = (!dna || dna >= maxLast) ? dna : MBO_DNA_NOT_WELLFORMED; /* :264 */
        if (dna == MBO_DNA_NOT_WELLFORMED) { *status = -4; break; }
        if (c == '>') last_char = c; /* :266; otherwise it keeps the '>' of this record and the next call ends at :244 */
        if (nrec < recCap) {
            recs[nrec].hdrOff = hdrOff;
            recs[nrec].hdrLen = hdrEnd - hdrOff;
            recs[nrec].seqOff = outl;
            recs[nrec].seqLen = seql;
        }
        nrec++;
        outl += seql;
    }
    *dnaLineLenOut = dna;
    *maxLastOut = maxLast;
    return nrec;
}

/* =====================================================================================================================
 * Row a14: the two legacy symbol packers.  Restated from /root/reference/coders/SymbolsPackingFacility.cpp and
 * coders/VarLenDNACoder.cpp; pinned against the reference's own classes by tests/test_coders_oracle.py
 * (oracle/_ref/libref_coders.so) and golden digests.
 * ===================================================================================================================== */

/* SymbolsPackingFacility::maxSymbolsPerElement, SymbolsPackingFacility.cpp:102-111 */
int mbo_symbols_per_element(int nSymbols) {
    int k = 0;
    uint64_t pw = 1;
    while (pw * (uint64_t) nSymbols <= 256) {
        pw *= (uint64_t) nSymbols;
        k++;
    }
    return k;
}

/* packSequence :123-139 with packSymbols :161-166 (= the positional value for symbols of the alphabet) and
 * packSuffixSymbols :149-159.  Returns the packed length, or UINT64_MAX for a symbol outside the alphabet. */
uint64_t mbo_symbols_pack(const char* symbols, int nSymbols, const uint8_t* src, uint64_t n, uint8_t* dst) {
    int order[256];
    for (int i = 0; i < 256; i++) order[i] = -1;
    for (int i = 0; i < nSymbols; i++) order[(uint8_t) symbols[i]] = i;
    const int per = mbo_symbols_per_element(nSymbols);
    uint64_t o = 0;
    for (uint64_t i = 0; i < n; i += (uint64_t) per) {
        unsigned v = 0;
        for (int j = 0; j < per; j++) {
            v *= (unsigned) nSymbols;
            if (i + (uint64_t) j < n) {
                const int k = order[src[i + (uint64_t) j]];
                if (k < 0) return UINT64_MAX;
                v += (unsigned) k;
            }
        }
        dst[o++] = (uint8_t) v;
    }
    return o;
}

/* reverseSequence(sequence, pos, length, destPtr) :190-215 */
void mbo_symbols_unpack(const char* symbols, int nSymbols, const uint8_t* packed, uint64_t pos, uint64_t length, uint8_t* dst) {
    const int per = mbo_symbols_per_element(nSymbols);
    for (uint64_t k = 0; k < length; k++) {
        const uint64_t p = pos + k;
        unsigned v = packed[p / (uint64_t) per];
        const int j = (int) (p % (uint64_t) per);
        for (int t = per - 1; t > j; t--) v /= (unsigned) nSymbols;
        dst[k] = (uint8_t) symbols[v % (unsigned) nSymbols];
    }
}

static uint32_t mbo_vl_idx(uint32_t v) { /* getLUTidx, VarLenDNACoder.h:40-43 */
    return (v & 0x3Fu) | ((v & 0x3F00u) >> 2) | ((v & 0x3F0000u) >> 4) | ((v & 0x3F000000u) >> 6);
}

/* initUsing :10-36 -> book[256][5] and the 2^24-entry table (caller frees *lutOut).  Returns the number of words or -1. */
static int mbo_vl_init(const char* codes, char book[256][5], uint8_t** lutOut) {
    memset(book, 0, 256 * 5);
    int counter = 0;
    const char* p = codes;
    for (;;) {
        const char* nl = strchr(p, '\n');
        const size_t len = nl ? (size_t) (nl - p) : strlen(p);
        if (counter > 255 || len > 4) return -1;
        memcpy(book[counter++], p, len);
        if (!nl) break;
        p = nl + 1;
    }
    if (strlen(book[0]) != 1) return -1;
    uint8_t* lut = (uint8_t*) calloc(1u << 24, 1);
    for (int i = 0; i < counter; i++) {
        uint32_t t = 0;
        memcpy(&t, book[i], 4);
        lut[mbo_vl_idx(t)] = (uint8_t) i;
    }
    *lutOut = lut;
    return counter;
}

/* encode :55-112.  Returns the number of codes, UINT64_MAX for an unexpected symbol or a bad book.  n < 4: the tail loop
 * only (the reference's `pos <= srcLen - 4` wraps around there). */
uint64_t mbo_varlen_encode(const char* codes, const uint8_t* src, uint64_t n, uint8_t* dst) {
    char book[256][5];
    uint8_t* lut = NULL;
    if (mbo_vl_init(codes, book, &lut) < 0) return UINT64_MAX;
    uint64_t out = 0, pos = 0;
    while (n >= 4 && pos <= n - 4) {
        uint32_t t = 0;
        memcpy(&t, src + pos, 4);
        t = mbo_vl_idx(t);
        if (lut[t] != 0) { dst[out++] = lut[t]; pos += 4; continue; }
        t &= 0x0003FFFFu;
        if (lut[t] != 0) { dst[out++] = lut[t]; pos += 3; continue; }
        t &= 0x00000FFFu;
        if (lut[t] != 0) { dst[out++] = lut[t]; pos += 2; continue; }
        t &= 0x0000003Fu;
        if (lut[t] == 0 && t != (65u & 0x3Fu)) { free(lut); return UINT64_MAX; }
        dst[out++] = lut[t];
        pos++;
    }
    while (pos < n) {
        uint32_t t = 0;
        memcpy(&t, src + pos, (size_t) (n - pos));
        t = mbo_vl_idx(t);
        t &= 0x0003FFFFu;
        if (n - pos >= 3 && lut[t] != 0) { dst[out++] = lut[t]; pos += 3; continue; }
        t &= 0x00000FFFu;
        if (n - pos >= 2 && lut[t] != 0) { dst[out++] = lut[t]; pos += 2; continue; }
        t &= 0x0000003Fu;
        if (lut[t] == 0 && t != (65u & 0x3Fu)) { free(lut); return UINT64_MAX; }
        dst[out++] = lut[t];
        pos++;
    }
    free(lut);
    return out;
}

/* decode :114-128.  Returns the decoded length (the caller compares it with the expected one), UINT64_MAX for a bad book. */
uint64_t mbo_varlen_decode(const char* codes, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dstCap) {
    char book[256][5];
    uint8_t* lut = NULL;
    if (mbo_vl_init(codes, book, &lut) < 0) return UINT64_MAX;
    free(lut);
    uint64_t out = 0;
    for (uint64_t i = 0; i < n; i++)
        for (const char* w = book[src[i]]; *w; w++) {
            if (out < dstCap) dst[out] = (uint8_t) *w;
            out++;
        }
    return out;
}

/* =====================================================================================================================
 * Row f3: the reverse-complement redundancy pass of mode 3.  SimpleSequenceMatcher::rcMatchSequence
 * (/root/reference/matching/SimpleSequenceMatcher.cpp:165-176 -> markAndRemoveExactMatches :69-148 ->
 * exactMatchSequence :24-56, resolveMappingCollisionsInTheSameText :150-163) over CopMEMMatcher
 * (matching/copmem/CopMEMMatcher.cpp: initParams :63-93, calcCoprimes :108-137, genCumm / processRef :139-227 -- the
 * single-threaded build, i.e. `-t 1` -- and processExactMatchQueryTight :332-484) with the hash
 * maRushPrime1HashSparsified (utils/Hashes.h:43-66).  32-bit case only (sequence < 4 GiB).
 * The l / r 4-byte prefilter of the query loop (:388-393, :449-452) is left out: a match of at least minMatchLength that
 * contains the K-mer extends by LK2 = (L - K) / 2 on one side, so the prefilter never rejects an accepted match, and a
 * candidate it lets through wrongly fails the length / K-mer test anyway (pinned against the reference's own classes,
 * tests/test_rcmatch_oracle.py).
 * ===================================================================================================================== */

static uint32_t mbo_hash_sparsified(const uint8_t* s, int K) { /* utils/Hashes.h:43-66 */
    uint64_t h = (uint64_t) K;
    uint32_t j = 0, k;
    while (j < 3) {
        memcpy(&k, s, 4);
        k &= 0x00FFFFFFu;
        k += j++;
        h ^= k;
        h *= 171717;
        s += 4;
    }
    while (j < (uint32_t) K / 4) {
        memcpy(&k, s, 4);
        k &= 0x0000FFFFu;
        k += j++;
        h ^= k;
        h *= 171717;
        s += 4;
    }
    return (uint32_t) h;
}

/* PgHelpers::reverseComplement (utils/helper.cpp:412-417) with the case-preserving table of :312-358 */
static void mbo_complement_lut(uint8_t lut[256]) {
    for (int i = 0; i < 256; i++) lut[i] = (uint8_t) i;
    lut[0x7F] = 0; /* the reference fills CHAR_MIN .. CHAR_MAX - 1 only */
    const char* from = "ACGTNYRKMBDHVWS";
    const char* to = "TGCANRYMKVHDBSW";
    for (int i = 0; from[i]; i++) {
        lut[(uint8_t) from[i]] = (uint8_t) to[i];
        lut[(uint8_t) (from[i] + 32)] = (uint8_t) (to[i] + 32);
    }
    lut['U'] = 'U';
    lut['u'] = 'u';
}

void mbo_reverse_complement(const uint8_t* src, uint64_t n, uint8_t* dst) {
    uint8_t lut[256];
    mbo_complement_lut(lut);
    for (uint64_t i = 0; i < n; i++) dst[n - 1 - i] = lut[src[i]];
}

/* CopMEMMatcher::initParams / calcCoprimes.  Returns 0, or -1 for "Minimal matching length too short" / L-K mismatch. */
int mbo_copmem_params(uint64_t N, uint32_t L, uint32_t minMatchLength, int* K, int* k1, int* k2, uint32_t* hashSize) {
    if (minMatchLength > L) minMatchLength = L;
    int k;
    if (L > 110) k = 56;
    else if (L > 62) k = 44;
    else if (L > 53) k = 40;
    else if (L > 46) k = 36;
    else if (L > 42) k = 32;
    else if (L > 32) k = 28;
    else k = ((int) L / 4 - 1) * 4;
    if (minMatchLength < 24) return -1;
    const int kmm = ((int) minMatchLength / 4 - 1) * 4;
    if (kmm < k) k = kmm;
    const int t = (int) L - k + 1;
    if (t <= 0) return -1;
    if (t >= 20) {
        *k1 = (int) pow((double) t, 0.5) + 1;
        *k2 = *k1 - 1;
        if (*k1 * *k2 > t) { --*k2; --*k1; }
    } else if (t >= 15) { *k1 = 5; *k2 = 3; }
    else if (t >= 12) { *k1 = 4; *k2 = 3; }
    else if (t >= 10) { *k1 = 5; *k2 = 2; }
    else if (t >= 6) { *k1 = 3; *k2 = 2; }
    else { *k1 = t; *k2 = 1; }
    *K = k;
    uint8_t i = 24;
    uint32_t hs;
    do {
        hs = ((uint32_t) 1) << (i++);
    } while (i <= 31 && hs < N / (uint64_t) *k1);
    *hashSize = hs;
    return 0;
}

static int mbo_rcm_cmp(const void* a, const void* b) { /* TextMatch::operator< (matching/TextMatchers.h:30-40): dest, src, length */
    const mbo_rcm_t* x = (const mbo_rcm_t*) a;
    const mbo_rcm_t* y = (const mbo_rcm_t*) b;
    if (x->dst != y->dst) return x->dst < y->dst ? -1 : 1;
    if (x->src != y->src) return x->src < y->src ? -1 : 1;
    if (x->len != y->len) return x->len < y->len ? -1 : 1;
    return 0;
}

static uint64_t mbo_frugal_bytes(uint8_t* dst, uint64_t at, uint64_t v) { /* writeUIntByteFrugal, utils/helper.cpp:217-225 */
    while (v >= 128) {
        dst[at++] = (uint8_t) (128 + v % 128);
        v /= 128;
    }
    dst[at++] = (uint8_t) v;
    return at;
}

/* The raw match list of CopMEMMatcher::matchTexts(resMatches, reverseComplement(seq), destIsSrc = true, revComplMatching =
 * true, minMatchLength) as the query loop pushes it (positions in the reverse-complemented text).  Caller frees *out. */
int64_t mbo_copmem_rc_matches(const uint8_t* seq, uint64_t N, uint32_t L, uint32_t minMatchLength, mbo_rcm_t** out) {
    *out = NULL;
    if (N >= (1ull << 32)) return -2;
    int K, k1, k2;
    uint32_t hashSize;
    if (minMatchLength > L) minMatchLength = L;
    if (mbo_copmem_params(N, L, minMatchLength, &K, &k1, &k2, &hashSize) < 0) return -1;
    if (minMatchLength < (uint32_t) K) return -1;
    const uint32_t mask = hashSize - 1;
    /* ---- genCumm + processRef, single-threaded (:139-227) ---- */
    uint32_t* cumm = (uint32_t*) calloc((size_t) hashSize + 2, 4);
    uint8_t* skipped = (uint8_t*) calloc((size_t) (N / (uint64_t) k1 + 2), 1);
    for (uint64_t i = 0, s = 0; i + (uint64_t) K <= N; i += (uint64_t) k1, s++) {
        const uint32_t h = (mbo_hash_sparsified(seq + i, K) & mask) + 2;
        if (cumm[h] <= 12) ++cumm[h];
        else skipped[s] = 1;
    }
    for (uint64_t x = 1; x < (uint64_t) hashSize + 2; x++) cumm[x] += cumm[x - 1];
    const uint32_t hashCount = cumm[hashSize + 1];
    uint32_t* sampled = (uint32_t*) malloc(((size_t) hashCount + 2) * 4);
    for (uint64_t i = 0, s = 0; i + (uint64_t) K <= N; i += (uint64_t) k1, s++) {
        if (skipped[s]) continue;
        const uint32_t h = (mbo_hash_sparsified(seq + i, K) & mask) + 1;
        sampled[cumm[h]++] = (uint32_t) i;
    }
    free(skipped);
    /* ---- the query: reverseComplement(seq) against the index of seq (:332-484) ---- */
    uint8_t* q = (uint8_t*) malloc((size_t) N + 8);
    mbo_reverse_complement(seq, N, q);
    const uint64_t N2 = N;
    const int skip = K / k1 - 1;
    const uint64_t MULTI = 256, k2MULTI = (uint64_t) k2 * MULTI;
    uint64_t cap = 1024, n = 0;
    mbo_rcm_t* res = (mbo_rcm_t*) malloc(cap * sizeof(mbo_rcm_t));
    /* one query position: returns 1 when the walk skips ahead (a push, or a seed inside the last match) */
#define MBO_RC_QUERY(POS, HIT)                                                                                          \
    do {                                                                                                                \
        const uint64_t pos_ = (POS);                                                                                    \
        const uint32_t h_ = mbo_hash_sparsified(q + pos_, K) & mask;                                                    \
        HIT = 0;                                                                                                        \
        for (uint32_t j = cumm[h_]; j < cumm[h_ + 1]; j++) {                                                            \
            const uint64_t src_ = sampled[j];                                                                           \
            if (N2 - src_ < pos_) continue;                                                                             \
            if (n > 0 && pos_ - src_ == res[n - 1].dst - res[n - 1].src && pos_ + (uint64_t) K < res[n - 1].dst + res[n - 1].len) { \
                HIT = 1;                                                                                                \
                break;                                                                                                  \
            }                                                                                                           \
            uint64_t p1 = src_ + (uint64_t) K - 1, p2 = pos_ + (uint64_t) K - 1;                                        \
            while (++p1 != N && ++p2 != N2 && seq[p1] == q[p2]) {}                                                      \
            const uint64_t right = p1;                                                                                  \
            int64_t a = (int64_t) src_, b = (int64_t) pos_;                                                             \
            while (a != 0 && b != 0 && seq[a] == q[b]) { a--; b--; }                                                    \
            if ((int64_t) right - a > (int64_t) minMatchLength && memcmp(seq + src_, q + pos_, (size_t) K) == 0) {      \
                if (n == cap) { cap *= 2; res = (mbo_rcm_t*) realloc(res, cap * sizeof(mbo_rcm_t)); }                   \
                res[n].src = (uint64_t) (a + 1);                                                                        \
                res[n].len = (uint64_t) ((int64_t) right - a - 1);                                                      \
                res[n].dst = (uint64_t) (b + 1);                                                                        \
                n++;                                                                                                    \
                HIT = 1;                                                                                                \
                break;                                                                                                  \
            }                                                                                                           \
        }                                                                                                               \
    } while (0)
    uint64_t i1 = 0;
    for (i1 = 0; i1 + (uint64_t) K + k2MULTI < N2 + 1; i1 += k2MULTI) {
        for (uint64_t i2 = 0; i2 < MULTI; ++i2) {
            int hit;
            MBO_RC_QUERY(i1 + i2 * (uint64_t) k2, hit);
            if (hit) i2 += (uint64_t) skip;
        }
    }
    for (; i1 + (uint64_t) K < N2 + 1; i1 += (uint64_t) k2) {
        int hit;
        MBO_RC_QUERY(i1, hit);
        if (hit) i1 += (uint64_t) skip * (uint64_t) k2;
    }
#undef MBO_RC_QUERY
    free(q);
    free(sampled);
    free(cumm);
    *out = res;
    return (int64_t) n;
}

/* rcMatchSequence.  seq (n bytes) is rewritten in place; returns its new length.  mapOff needs 4 bytes per match, mapLen
 * up to 10 bytes per match + 10.  Returns UINT64_MAX on a parameter error (the reference exits). */
uint64_t mbo_rc_match_sequence(uint8_t* seq, uint64_t N, uint32_t targetMatchLength, uint32_t minMatchLength, uint8_t* mapOff,
                               uint64_t* mapOffLen, uint8_t* mapLen, uint64_t* mapLenLen, uint64_t* nMatchesRaw) {
    *mapOffLen = *mapLenLen = 0;
    if (nMatchesRaw) *nMatchesRaw = 0;
    if (N < targetMatchLength) return N; /* no matcher: both maps stay empty (SimpleSequenceMatcher.cpp:17-19, 72-76) */
    mbo_rcm_t* m = NULL;
    const int64_t nm = mbo_copmem_rc_matches(seq, N, targetMatchLength, minMatchLength, &m);
    if (nm < 0) return UINT64_MAX;
    if (minMatchLength == UINT32_MAX) minMatchLength = targetMatchLength;
    if (nMatchesRaw) *nMatchesRaw = (uint64_t) nm;
    uint64_t n = (uint64_t) nm;
    for (uint64_t i = 0; i < n; i++) { /* correctDestPositionDueToRevComplMatching :58-61, resolveMappingCollisions :150-163 */
        m[i].dst = N - (m[i].dst + m[i].len);
        if (m[i].src > m[i].dst) {
            const uint64_t t = m[i].src;
            m[i].src = m[i].dst;
            m[i].dst = t;
        }
        if (m[i].src + m[i].len > m[i].dst) {
            const uint64_t margin = (m[i].src + m[i].len - m[i].dst + 1) / 2;
            m[i].len -= margin;
            m[i].dst += margin;
        }
    }
    uint64_t lo = mbo_frugal_bytes(mapLen, 0, minMatchLength);
    qsort(m, (size_t) n, sizeof(mbo_rcm_t), mbo_rcm_cmp);
    uint64_t u = 0;
    for (uint64_t i = 0; i < n; i++)
        if (u == 0 || mbo_rcm_cmp(&m[i], &m[u - 1]) != 0) m[u++] = m[i];
    n = u;
    uint64_t pos = 0, nPos = 0, oo = 0;
    for (uint64_t i = 0; i < n; i++) { /* :99-125 */
        mbo_rcm_t* t = &m[i];
        if (t->dst < pos) {
            const uint64_t overflow = pos - t->dst;
            if (overflow >= t->len) {
                t->len = 0;
                continue;
            }
            t->len -= overflow;
            t->dst += overflow; /* revComplMatching: posSrcText stays */
        }
        if (t->len < minMatchLength) continue;
        const uint64_t length = t->dst - pos;
        memmove(seq + nPos, seq + pos, (size_t) length);
        nPos += length;
        seq[nPos++] = (uint8_t) ('$' + 128); /* RC_MATCH_MARK, mbgccoder/MBGC_Params.h:48 */
        const uint32_t src32 = (uint32_t) t->src;
        memcpy(mapOff + oo, &src32, 4);
        oo += 4;
        lo = mbo_frugal_bytes(mapLen, lo, t->len - minMatchLength);
        pos = t->dst + t->len;
    }
    memmove(seq + nPos, seq + pos, (size_t) (N - pos));
    nPos += N - pos;
    free(m);
    *mapOffLen = oo;
    *mapLenLen = lo;
    return nPos;
}
