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
