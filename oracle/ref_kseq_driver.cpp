// oracle/ref_kseq_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" driver over the REFERENCE's own FASTA reader: /root/reference/utils/kseq.h is included from where
// it lies and instantiated over an in-memory file, exactly as matching/MultipleGenomeMatchingProcessor.cpp:7 does over
// its file type (KSEQ_INIT(mgmpInFile, mgmpInRead)).  oracle/Makefile builds it into oracle/_ref/libref_kseq.so; the C
// restatement (oracle/mbgc_oracle.c: mbo_fasta_parse) and the CUDA ingest path are pinned against it.  No algorithm
// of its own: every record comes from kseq_read_lossless_fasta (utils/kseq.h:233-274).
#include <cstdint>
#include <cstring>

#include "utils/kseq.h"

namespace {
struct MemFile {
    const uint8_t* p;
    uint64_t n, at;
};
int mem_read(MemFile* f, unsigned char* buf, int len) {
    const uint64_t left = f->n - f->at;
    const int k = left < static_cast<uint64_t>(len) ? static_cast<int>(left) : len;
    memcpy(buf, f->p + f->at, k);
    f->at += k;
    return k;
}
}  // namespace

KSEQ_INIT(MemFile*, mem_read)

extern "C" {

// same contract as mbo_fasta_parse (oracle/mbgc_oracle.h); header offsets are recovered from the stream position
uint64_t refk_fasta_parse(const uint8_t* buf, uint64_t n, uint64_t* recs /* 4 per record: hdrLen, seqOff, seqLen, 0 */,
                          uint64_t recCap, uint8_t* seqOut, uint8_t* hdrOut /* headers concatenated */, uint64_t* dnaLineLen,
                          uint64_t* maxLastDnaLineLen, int* status) {
    MemFile f{buf, n, 0};
    kseq_t* seq = kseq_init(&f);
    uint64_t nrec = 0, outl = 0, hdrl = 0;
    int64_t r;
    while ((r = kseq_read_lossless_fasta(seq)) >= 0) {
        if (nrec < recCap) {
            recs[4 * nrec + 0] = seq->name.l;
            recs[4 * nrec + 1] = outl;
            recs[4 * nrec + 2] = seq->seq.l;
            recs[4 * nrec + 3] = 0;
        }
        memcpy(hdrOut + hdrl, seq->name.s, seq->name.l);
        hdrl += seq->name.l;
        memcpy(seqOut + outl, seq->seq.s, seq->seq.l);
        outl += seq->seq.l;
        nrec++;
    }
    *status = static_cast<int>(r);
    *dnaLineLen = seq->dnaLineLen;
    *maxLastDnaLineLen = seq->maxLastDnaLineLen;
    kseq_destroy(seq);
    return nrec;
}

}  // extern "C"
