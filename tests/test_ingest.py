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
