"""Host-side chop output (dcb200_chop_write_bgzf: record assembly + BGZF on host threads; no GPU needed): the
decompressed bytes must equal a plain sequential assembly of the same records for any thread count, and the file must
be a well-formed BGZF stream (htslib block layout + EOF block)."""
import gzip
import struct

import numpy as np
import pytest

from deepchopper_b200 import synth
from deepchopper_b200.chop import write_chopped_fastq
from deepchopper_b200.encode import index_fastq

P, T_, I_, AD, UN = 0, 1, 2, 3, 4


def _case(rng, n_reads, max_iv=4):
    lens = rng.integers(30, 4000, n_reads)
    recs = synth.fastq_reads(rng, n_reads, lengths=lens)
    recs = [(rid + (" desc=%d extra" % i if i % 5 == 0 else ""), s, q) for i, (rid, s, q) in enumerate(recs)]
    buf = np.frombuffer(synth.fastq_text(recs), dtype=np.uint8)
    ix = index_fastq(buf)
    R = len(ix)
    has = (rng.random(R) > 0.1).astype(np.uint8)
    act = rng.integers(0, 5, R).astype(np.uint8)
    n_ad = rng.integers(0, max_iv + 1, R).astype(np.int32)
    n_keep = rng.integers(0, max_iv + 2, R).astype(np.int32)
    ad = np.zeros((R, max_iv, 2), np.int32)
    keep = np.zeros((R, max_iv + 1, 2), np.int32)
    pseqs = []
    for r in range(R):
        n = int(lens[r])
        plen = n if r % 7 else max(1, n - 9)            # a prediction shorter than the FASTQ read (truncated)
        pseqs.append(np.frombuffer(bytes(rng.choice(list(b"ACGTN"), plen).astype(np.uint8)), dtype=np.uint8).copy())
        for arr, cnt in ((ad, n_ad), (keep, n_keep)):
            cuts = np.sort(rng.integers(0, n + 20, 2 * arr.shape[1]))    # some intervals reach past the end: clamped
            arr[r] = cuts.reshape(-1, 2)
    ptr = np.array([p.ctypes.data for p in pseqs], dtype=np.uint64)
    plen = np.array([p.size for p in pseqs], dtype=np.int32)
    return recs, ix, has, pseqs, ptr, plen, act, n_ad, ad, n_keep, keep


def _expected(recs, has, pseqs, act, n_ad, ad, n_keep, keep):
    out, nrec = [], 0
    for r, (head, seq, qual) in enumerate(recs):
        if not has[r]:
            continue
        rid = head.split(" ")[0]
        ps = pseqs[r].tobytes().decode()
        if act[r] == P:
            out.append(f"@{head}\n{seq}\n+\n{qual}\n")
            nrec += 1
        elif act[r] == UN:
            out.append(f"@{rid}\n{ps}\n+\n{qual}\n")
            nrec += 1
        elif act[r] == AD:
            for s, e in ad[r, :n_ad[r]]:
                out.append(f"@{rid}|{s}:{e}\n{ps[s:e]}\n+\n{qual[s:e]}\n")
                nrec += 1
        else:
            tag = "T" if act[r] == T_ else "I"
            for s, e in keep[r, :n_keep[r]]:
                out.append(f"@{rid}|{s}:{e}|{tag}\n{ps[s:e]}\n+\n{qual[s:e]}\n")
                nrec += 1
    return "".join(out), nrec


def _check_bgzf(raw: bytes):
    pos, nblocks = 0, 0
    while pos < len(raw):
        assert raw[pos:pos + 4] == b"\x1f\x8b\x08\x04"
        xlen = struct.unpack_from("<H", raw, pos + 10)[0]
        assert xlen == 6 and raw[pos + 12:pos + 16] == b"BC\x02\x00"
        bsize = struct.unpack_from("<H", raw, pos + 16)[0] + 1
        isize = struct.unpack_from("<I", raw, pos + bsize - 4)[0]
        assert isize <= 0xff00 and bsize <= 65536
        pos += bsize
        nblocks += 1
    assert pos == len(raw)
    assert raw[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    return nblocks


@pytest.mark.parametrize("threads,level", [(1, 6), (3, 6), (8, 1), (4, 0)])
def test_writer_matches_sequential_assembly(tmp_path, threads, level):
    rng = np.random.default_rng(7)
    recs, ix, has, pseqs, ptr, plen, act, n_ad, ad, n_keep, keep = _case(rng, 700)
    path = str(tmp_path / f"out{threads}_{level}.fq.gz")
    nrec, ntext = write_chopped_fastq(path, ix, has, ptr, plen, act, n_ad, ad, n_keep, keep, threads=threads,
                                        level=level)
    want, want_rec = _expected(recs, has, pseqs, act, n_ad, ad, n_keep, keep)
    raw = open(path, "rb").read()
    assert _check_bgzf(raw) >= 2
    got = gzip.decompress(raw).decode()
    assert nrec == want_rec and ntext == len(want)
    assert got == want


def test_writer_empty_and_all_dropped(tmp_path):
    rng = np.random.default_rng(8)
    recs, ix, has, pseqs, ptr, plen, act, n_ad, ad, n_keep, keep = _case(rng, 12)
    has[:] = 0
    path = str(tmp_path / "none.fq.gz")
    assert write_chopped_fastq(path, ix, has, ptr, plen, act, n_ad, ad, n_keep, keep, threads=4) == (0, 0)
    raw = open(path, "rb").read()
    assert _check_bgzf(raw) == 1 and gzip.decompress(raw) == b""


def test_writer_rejects_unknown_action(tmp_path):
    from deepchopper_b200._native import Dcb200Error
    rng = np.random.default_rng(9)
    recs, ix, has, pseqs, ptr, plen, act, n_ad, ad, n_keep, keep = _case(rng, 5)
    has[:] = 1
    act[2] = 9
    with pytest.raises(Dcb200Error):
        write_chopped_fastq(str(tmp_path / "bad.fq.gz"), ix, has, ptr, plen, act, n_ad, ad, n_keep, keep, threads=2)


def test_writer_parts_append_to_one_valid_file(tmp_path):
    """dcb200_chop_write_bgzf_part: the records written in three appended parts (the last one closing the file with the
    BGZF end-of-file block) decompress to the one-call output, and the file is still a well-formed BGZF chain."""
    from deepchopper_b200.chop import WRITE_APPEND, WRITE_NO_EOF
    from deepchopper_b200.encode import FastqIndex
    rng = np.random.default_rng(10)
    recs, ix, has, pseqs, ptr, plen, act, n_ad, ad, n_keep, keep = _case(rng, 300)
    whole = str(tmp_path / "whole.fq.gz")
    nrec, ntext = write_chopped_fastq(whole, ix, has, ptr, plen, act, n_ad, ad, n_keep, keep, threads=3)
    parts = str(tmp_path / "parts.fq.gz")
    open(parts, "wb").close()
    tot_rec = tot_text = 0
    cuts = [0, 100, 101, 300]
    for a, b in zip(cuts[:-1], cuts[1:]):
        sub = FastqIndex(ix.buf, *[getattr(ix, f)[a:b] for f in ("name_off", "name_len", "head_len", "seq_off", "seq_len",
                                                                   "qual_off", "qual_len")])
        r, t = write_chopped_fastq(parts, sub, has[a:b], ptr[a:b], plen[a:b], act[a:b], n_ad[a:b], ad[a:b], n_keep[a:b],
                                   keep[a:b], threads=2, flags=WRITE_APPEND | (WRITE_NO_EOF if b < 300 else 0))
        tot_rec += r
        tot_text += t
    assert (tot_rec, tot_text) == (nrec, ntext)
    raw = open(parts, "rb").read()
    _check_bgzf(raw)
    assert gzip.decompress(raw) == gzip.decompress(open(whole, "rb").read())


def test_fastq_chunk_reader(tmp_path):
    """iter_fastq_chunks: pieces of whole records whose concatenation is the file, for plain and gzip input and for a
    file without trailing newline."""
    from deepchopper_b200.chop import iter_fastq_chunks
    rng = np.random.default_rng(0)
    text = synth.fastq_text(synth.fastq_reads(rng, 500, 20, 400))
    for name, data, want in (("a.fq", text, text), ("b.fq", text[:-1], text[:-1]), ("c.fq.gz", gzip.compress(text), text)):
        p = tmp_path / name
        p.write_bytes(data)
        for cb in (1000, 7777, 1 << 20):
            pieces = list(iter_fastq_chunks(str(p), cb))
            assert b"".join(x.tobytes() for x in pieces) == want
            assert sum(len(index_fastq(x)) for x in pieces) == 500
