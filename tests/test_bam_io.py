"""Host-side BAM / FASTA reading (clair3_b200/bam_io.py): the decoded-record arrays the GPU feature counter takes, out of a real
BGZF-compressed BAM file - round trips through the module's own writer, container checks against the BAM specification, and the
region fetch feeding the oracle the same counts as the whole file."""
import gzip
import os
import struct
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from clair3_b200 import bam_io, synth_reads as sr
from oracle import pileup_oracle as po

FIELDS = ("pos", "flag", "mapq", "l_qseq", "cigar_off", "cigar", "seq_off", "seq")


def test_bam_round_trip_and_container(tmp_path):
    rec, ref, rs = sr.random_alignment(3000, depth=15, read_len=400, seed=4, wild=True)
    path = str(tmp_path / "x.bam")
    refs = [("chrA", 1000), ("chr20", rs + len(ref))]
    bam_io.write_bam(path, rec, refs, tid=1)
    raw = open(path, "rb").read()
    # BGZF: every block is a gzip member with the 'BC' extra subfield carrying its size; the file ends with the 28-byte EOF marker
    off, blocks = 0, 0
    while off < len(raw):
        assert raw[off:off + 4] == b"\x1f\x8b\x08\x04" and raw[off + 12:off + 16] == b"BC\x02\x00"
        bsize, = struct.unpack_from("<H", raw, off + 16)
        off += bsize + 1
        blocks += 1
    assert off == len(raw) and blocks >= 3 and raw.endswith(bam_io._EOF)
    plain = gzip.decompress(raw)
    assert plain[:4] == b"BAM\x01"
    got, refs2 = bam_io.read_bam(path, with_names=True)
    assert refs2 == refs and (got["tid"] == 1).all() and got["names"][0] == "r0"
    for k in FIELDS:
        assert got[k].dtype == rec[k].dtype and np.array_equal(got[k], rec[k]), k
    # the bin of every record follows the specification's reg2bin
    _, first = bam_io.read_header(plain)
    tid, p, l_name, mq, bin_, n_cig, fl, l_seq, *_ = bam_io._CORE.unpack_from(plain, first + 4)
    span = int((rec["cigar"][:n_cig] >> 4)[np.isin(rec["cigar"][:n_cig] & 15, (0, 2, 3, 7, 8))].sum())
    assert bin_ == bam_io._reg2bin(p, p + max(span, 1)) and p == rec["pos"][0]


def test_region_fetch_feeds_the_counter_the_same_counts(tmp_path):
    rec, ref, rs = sr.random_alignment(6000, depth=20, read_len=500, seed=6)
    path, fa = str(tmp_path / "y.bam"), str(tmp_path / "ref.fa")
    bam_io.write_bam(path, rec, [("chr20", rs + len(ref))])
    bam_io.write_fasta(fa, [("chrOther", "ACGT" * 50), ("chr20", "N" * rs + ref)])
    sub, _ = bam_io.read_bam(path, "chr20", 3000, 4500)
    assert 0 < len(sub["pos"]) < len(rec["pos"])
    assert sub["pos"].min() < 3000 <= (sub["pos"] + 1).max() and sub["pos"].max() < 4500
    ref_start = 3000 - 1000
    bases = bam_io.read_fasta(fa, "chr20", ref_start, 4500 + 1000)              # the reference's +-1000 window (src/clair3_pileup.c:184-186)
    assert bases == ("N" * rs + ref)[ref_start:5500]
    whole = po.clair3_pileup(rec, 3000, 4500, ref, rs, alt_info=True)
    part = po.clair3_pileup(sub, 3000, 4500, bases, ref_start, alt_info=True)
    for k in ("matrix", "major", "stats", "cand_cols", "cand_ok"):
        assert np.array_equal(whole[k], part[k]), k
    assert whole["alt_info"] == part["alt_info"] and len(part["alt_info"]) > 3
    with pytest.raises(ValueError, match="not in the BAM header"):
        bam_io.read_bam(path, "chrX")
    with pytest.raises(ValueError, match="not found"):
        bam_io.read_fasta(fa, "chrX")


def test_records_validate_as_counter_input(tmp_path):
    from clair3_b200 import pileup_counts as pc
    rec, ref, rs = sr.random_alignment(500, depth=5, read_len=100, seed=8)
    path = str(tmp_path / "z.bam")
    bam_io.write_bam(path, rec, [("c", 10000)])
    got, _ = bam_io.read_bam(path, "c")
    b = pc.BamRecords.from_dict(got)
    assert b.n_reads == len(rec["pos"]) and b.nbytes() > 0
    with pytest.raises(ValueError, match="not a BAM"):
        bam_io.read_header(b"nope")
