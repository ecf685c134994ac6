"""File-level FASTQ ingest (dcb200_read_file_inflate; host threads, no GPU needed): plain, gzip and BGZF files give the
same bytes as Python's gzip module, for any thread count; corrupt input is an error."""
import gzip

import numpy as np
import pytest

from deepchopper_b200 import synth
from deepchopper_b200._native import Dcb200Error
from deepchopper_b200.chop import write_chopped_fastq
from deepchopper_b200.encode import index_fastq, read_fastq_bytes


def _text(n=400, seed=3):
    rng = np.random.default_rng(seed)
    return synth.fastq_text(synth.fastq_reads(rng, n, lengths=rng.integers(50, 5000, n)))


@pytest.mark.parametrize("threads", [1, 4, 0])
def test_plain_gzip_bgzf_roundtrip(tmp_path, threads):
    text = _text()
    (tmp_path / "a.fq").write_bytes(text)
    with gzip.open(tmp_path / "a.fq.gz", "wb") as f:
        f.write(text)
    with open(tmp_path / "two.fq.gz", "wb") as f:        # two concatenated gzip members
        f.write(gzip.compress(text[:1000]) + gzip.compress(text[1000:]))
    # BGZF written by the native chop writer: every record passes through verbatim
    ix = index_fastq(np.frombuffer(text, dtype=np.uint8))
    R = len(ix)
    z = np.zeros
    write_chopped_fastq(str(tmp_path / "b.fq.gz"), ix, np.ones(R, np.uint8), z(R, np.uint64), z(R, np.int32), z(R, np.uint8),
                        z(R, np.int32), z((R, 1, 2), np.int32), z(R, np.int32), z((R, 2, 2), np.int32), threads=3)
    for name in ("a.fq", "a.fq.gz", "two.fq.gz", "b.fq.gz"):
        got = read_fastq_bytes(str(tmp_path / name), threads=threads)
        assert got.dtype == np.uint8 and got.tobytes() == text, name
    assert gzip.open(tmp_path / "b.fq.gz", "rb").read() == text


def test_empty_and_corrupt(tmp_path):
    (tmp_path / "e.fq").write_bytes(b"")
    assert read_fastq_bytes(str(tmp_path / "e.fq")).size == 0
    text = _text(50)
    raw = bytearray(gzip.compress(text))
    raw[len(raw) // 2] ^= 0x5a
    (tmp_path / "bad.fq.gz").write_bytes(bytes(raw))
    with pytest.raises(Dcb200Error):
        read_fastq_bytes(str(tmp_path / "bad.fq.gz"))
    (tmp_path / "cut.fq.gz").write_bytes(gzip.compress(text)[:-20])
    with pytest.raises(Dcb200Error):
        read_fastq_bytes(str(tmp_path / "cut.fq.gz"))
    with pytest.raises(Dcb200Error):
        read_fastq_bytes(str(tmp_path / "missing.fq"))


def test_native_index_equals_numpy_restatement():
    """dcb200_index_fastq (host threads) against the vectorised numpy restatement: random reads, CRLF line ends, tab / space
    descriptions, no trailing newline, trailing blank lines, 1 / 3 / all threads; the same inputs are rejected."""
    import pytest
    from deepchopper_b200 import synth
    from helpers_index import index_fastq_numpy
    rng = np.random.default_rng(11)
    recs = synth.fastq_reads(rng, 3000, 1, 400)
    recs[3] = (recs[3][0] + " some description", recs[3][1], recs[3][2])
    recs[4] = (recs[4][0] + "\ttabbed\tdescription", recs[4][1], recs[4][2])
    text = synth.fastq_text(recs)
    assert len(index_fastq(np.zeros(0, dtype=np.uint8))) == 0
    variants = [text, text[:-1], text + b"\n\n", text.replace(b"\n", b"\r\n"), b"\n"]
    for v in variants:
        buf = np.frombuffer(v, dtype=np.uint8)
        want = index_fastq_numpy(buf)
        for threads in (1, 3, 0):
            got = index_fastq(buf, threads)
            assert len(got) == len(want)
            for f in ("name_off", "name_len", "head_len", "seq_off", "seq_len", "qual_off", "qual_len"):
                assert np.array_equal(getattr(got, f), getattr(want, f)), (f, threads)
                assert getattr(got, f).dtype == getattr(want, f).dtype
    big = text * 40                       # several MB: more than one slice per thread
    g, w = index_fastq(np.frombuffer(big, dtype=np.uint8)), index_fastq_numpy(np.frombuffer(big, dtype=np.uint8))
    assert len(g) == 120000 and np.array_equal(g.seq_off, w.seq_off) and np.array_equal(g.name_len, w.name_len)
    for bad, msg in ((b"@a\nACGT\n+\nIII\n", "lengths differ"), (b"a\nACGT\n+\nIIII\n", "'@'"), (b"@a\nAC\n-\nII\n", "'\\+'"),
                     (b"@a\nAC\n+\n", "multiple of 4"), (b"@a\n\n+\n\n@b\nA\n+\nI\n", "empty sequence")):
        with pytest.raises(ValueError, match=msg):
            index_fastq(np.frombuffer(bad, dtype=np.uint8))
        with pytest.raises(ValueError):
            index_fastq_numpy(np.frombuffer(bad, dtype=np.uint8))
