"""Vectorised numpy restatement of the FASTQ record index (the checker of the native dcb200_index_fastq)."""
import numpy as np

from deepchopper_b200.encode import FastqIndex


def index_fastq_numpy(buf: np.ndarray) -> FastqIndex:
    """Vectorised newline index.  Validates like only_fq.py:38-85: '@' headers, equal seq/qual lengths."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    nl = np.flatnonzero(buf == 10)
    if buf.size and (nl.size == 0 or nl[-1] != buf.size - 1):
        nl = np.append(nl, buf.size)          # last line without trailing newline
    line_start = np.concatenate([[0], nl[:-1] + 1]).astype(np.int64)
    line_end = nl.astype(np.int64)
    # strip '\r'
    cr = (line_end > line_start) & (buf[np.maximum(line_end - 1, 0)] == 13)
    line_end = line_end - cr
    # drop trailing empty lines
    nlines = line_start.size
    while nlines and line_end[nlines - 1] == line_start[nlines - 1]:
        nlines -= 1
    if nlines % 4 != 0:
        raise ValueError(f"FASTQ has {nlines} lines, not a multiple of 4")
    R = nlines // 4
    hs, he = line_start[0:nlines:4], line_end[0:nlines:4]
    ss, se = line_start[1:nlines:4], line_end[1:nlines:4]
    ps = line_start[2:nlines:4]
    qs, qe = line_start[3:nlines:4], line_end[3:nlines:4]
    if R:
        if not (buf[hs] == ord("@")).all():
            raise ValueError("FASTQ record does not start with '@'")
        if not (buf[ps] == ord("+")).all():
            raise ValueError("FASTQ separator line does not start with '+'")
        if not ((se - ss) == (qe - qs)).all():
            bad = int(np.flatnonzero((se - ss) != (qe - qs))[0])
            raise ValueError(f"record {bad}: sequence and quality lengths differ")     # only_fq.py:49-57
        if ((se - ss) == 0).any():
            raise ValueError("empty sequence in FASTQ")                                # only_fq.py:44-47
    head_len = (he - hs - 1).astype(np.int32)
    # id = header up to the first blank
    name_len = head_len.copy()
    is_blank = (buf == 32) | (buf == 9)
    blank_pos = np.flatnonzero(is_blank)
    if blank_pos.size and R:
        j = np.searchsorted(blank_pos, hs + 1)
        has = j < blank_pos.size
        first = np.where(has, blank_pos[np.minimum(j, blank_pos.size - 1)], np.iinfo(np.int64).max)
        inside = first < he
        name_len = np.where(inside, first - hs - 1, head_len).astype(np.int32)
    return FastqIndex(buf, hs + 1, name_len, head_len, ss, (se - ss).astype(np.int32), qs, (qe - qs).astype(np.int32))


