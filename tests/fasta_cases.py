"""FASTA inputs shared by the reader tests (oracle vs the reference's kseq.h on the CPU, CUDA ingest vs oracle on the
GPU): well-formed files and every way utils/kseq.h:233-274 can end a file early."""
import os

import numpy as np

from util import FIXTURES, have_fixtures

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _seq(rng, n):
    return bytes(_ACGT[rng.integers(0, 4, n)])


def _fasta(records, width, trailing_newline=True, line_end=b"\n"):
    out = []
    for hdr, seq in records:
        out.append(b">" + hdr + line_end)
        for i in range(0, len(seq), width):
            out.append(seq[i:i + width] + line_end)
    data = b"".join(out)
    if not trailing_newline and data.endswith(line_end):
        data = data[:-len(line_end)]
    return data


def fasta_cases():
    rng = np.random.default_rng(99)
    recs = [(b"contig_%d some description %d" % (i, i), _seq(rng, n)) for i, n in enumerate((1000, 81, 80, 79, 1, 160, 0, 5000, 12))]
    cases = {
        "wellformed80": _fasta(recs, 80),
        "wellformed60_no_trailing_newline": _fasta(recs, 60, trailing_newline=False),
        "single_line_records": _fasta([(b"a", _seq(rng, 300)), (b"b", _seq(rng, 200)), (b"c", _seq(rng, 300))], 10**9),
        "single_line_then_wrapped_longer_last": _fasta([(b"a", _seq(rng, 100))], 10**9) + _fasta([(b"b", _seq(rng, 200))], 80),
        "wrapped_then_long_single_line": _fasta([(b"b", _seq(rng, 200))], 80) + _fasta([(b"a", _seq(rng, 100))], 10**9),
        "one_big_record": _fasta([(b"chromosome", _seq(rng, 300_007))], 70),
        "crlf": _fasta(recs[:4], 80, line_end=b"\r\n"),
        "gt_inside_line": b">h1\nACGT>ACGT\nACGTTGCA\nAC\n>h2\nAAAA\n",
        "lowercase_and_N": b">h1 x\nacgtnNNNNN\nACGTacgtRY\nAC\n>h2\nNNNNNNNNNN\n",
        "header_only_records": b">h1\n>h2\n>h3\nACGT\n>h4\n",
        "empty_header": b">\nACGT\nACGT\n>\nAC\n",
        "empty_file": b"",
        "only_gt": b">",
        "gt_and_newline": b">\n",
        "header_without_newline": b">header only",
        "not_fasta": b"ACGT\n>h\nACGT\n",
        "leading_newline": b"\n>h\nACGT\n",
        "empty_line_inside": b">h1\nACGT\n\nACGT\n>h2\nAC\n",
        "empty_line_at_end_of_record": b">h1\nACGT\nACGT\n\n>h2\nAC\n",
        "empty_line_after_header": b">h1\n\nACGT\n",
        "trailing_empty_line": b">h1\nACGT\nAC\n\n",
        "inconsistent_width_inside_record": b">h1\nACGTACGT\nACGTAC\nACGTACGT\nAC\n",
        "inconsistent_width_between_records": b">h1\nACGTACGT\nAC\n>h2\nACGTAC\nACGTAC\nA\n>h3\nAC\n",
        "last_line_longer": b">h1\nACGT\nACGTAC\n",
        "second_record_sets_smaller_width": b">h1\nACGTACGTAC\n>h2\nACGT\nACGT\nAC\n",
        "good_records_before_a_bad_one": _fasta(recs[:3], 80) + b">bad\nACGT\n\nAC\n" + _fasta(recs[3:5], 80),
        "many_small_records": _fasta([(b"r%d" % i, _seq(rng, int(rng.integers(0, 400)))) for i in range(500)], 80),
        "width_1": _fasta([(b"w", _seq(rng, 50))], 1),
        "binary_bytes": b">h\x00\xff\n\x00\x01\xfe\xff\n\x80\x81\x82\x83\nAB\n",
    }
    if have_fixtures():
        for f in sorted(os.listdir(FIXTURES)):
            if f.endswith(".fna"):
                cases["fixture_" + f] = open(os.path.join(FIXTURES, f), "rb").read()
    return cases
