"""CPU ORACLE (test infrastructure, NOT the product) for DeepChopper's smooth / interval / chop pass.

Literal restatement, in plain Python, of the reference's Rust code.  Every function cites the
reference file:line it follows (paths relative to /root/reference).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may
import this module.  The product path (``deepchopper_b200``) never does.

Parity status: PINNED for smoothing/chop by the reference's own Rust unit-test vectors
(tests/test_oracle_smooth.py replays every one of them) -- the Rust toolchain is absent in this
image so the reference binary itself cannot be run here.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

# src/default.rs:1-7
QUAL_OFFSET = 33
MIN_READ_LEN = 150
MIN_CHOPED_SEQ_LEN = 20
IGNORE_LABEL = -100

# src/smooth/utils.rs:6-25 -- token id -> base; anything else decodes to 'N' (:41-46)
ID_TABLE = {7: "A", 8: "C", 9: "G", 10: "T", 11: "N"}

Interval = Tuple[int, int]


def id_list2seq(ids: Sequence[int]) -> str:
    """src/smooth/utils.rs:34-46."""
    return "".join(ID_TABLE.get(int(i), "N") for i in ids)


def ascii_list2str(vals: Sequence[int]) -> str:
    """src/smooth/utils.rs:27-32."""
    return "".join(chr(int(v)) for v in vals)


def majority_voting(labels: Sequence[int], window_size: int) -> List[int]:
    """src/smooth/utils.rs:48-97.

    Even windows are bumped to odd (:50-54); the window is clipped on the left edge (:66) and
    shifted to full width on the right edge (:69-71); an exact two-way tie keeps the original
    label (:86-91).
    """
    if window_size % 2 == 0:
        window_size += 1
    half = window_size // 2
    n = len(labels)
    out = []
    for i in range(n):
        start = max(0, i - half)
        end = min(n, i + half + 1)
        if end == n and (end - start) < window_size:
            start = max(0, end - window_size)
        counts = {}
        for lab in labels[start:end]:
            counts[lab] = counts.get(lab, 0) + 1
        if len(counts) == 2:
            a, b = counts.values()
            if a == b:
                out.append(labels[i])
                continue
        # binary labels: no further ties possible (max_by_key's hash-order tie-break never fires)
        out.append(max(counts.items(), key=lambda kv: kv[1])[0])
    return out


def get_label_region(labels: Sequence[int]) -> List[Interval]:
    """src/utils.rs:671-695 -- including the ``start == 0`` sentinel quirk (a run that begins at
    index 0 loses its first base; a length-1 run at index 0 vanishes)."""
    regions = []
    start = 0
    end = 0
    for i, lab in enumerate(labels):
        if lab == 1:
            if start == 0:
                start = i
            end = i
        elif start != 0:
            regions.append((start, end + 1))
            start = 0
            end = 0
    if start != 0:
        regions.append((start, end + 1))
    return regions


def smooth_label_region(labels: Sequence[int], smooth_window_size: int, min_interval_size: int,
                        approved_interval_number: int) -> List[Interval]:
    """src/utils.rs:699-721 == Predict::smooth_and_select_intervals src/smooth/predict.rs:186-209."""
    regions = get_label_region(majority_voting(labels, smooth_window_size))
    results = [r for r in regions if r[1] - r[0] >= min_interval_size]
    if len(results) > approved_interval_number:
        return []
    return results


def summary_predict(predictions: Sequence[Sequence[int]], labels: Sequence[Sequence[int]],
                    ignore_label: int = IGNORE_LABEL):
    """src/utils.rs:9-55 -- keep positions whose label != ignore_label."""
    outp, outl = [], []
    for p, l in zip(predictions, labels):
        fp, fl = [], []
        for pv, lv in zip(p, l):
            if lv != ignore_label:
                fp.append(pv)
                fl.append(lv)
        outp.append(fp)
        outl.append(fl)
    return outp, outl


def generate_unmaped_intervals(intervals: Sequence[Interval], total_length: int) -> List[Interval]:
    """src/output/split.rs:260-292 -- the trailing kept piece stops at total_length-1 (:287-289)."""
    if not intervals:
        return [(0, total_length)]
    result = []
    current_start = 0
    for s, e in intervals:
        if current_start < s:
            result.append((current_start, s))
        current_start = e
    if current_start < total_length - 1:
        result.append((current_start, total_length - 1))
    return result


def remove_intervals_and_keep_left(seq, intervals: Sequence[Interval]):
    """src/output/split.rs:295-320."""
    ivs = sorted(intervals, key=lambda r: r[0])
    selected = generate_unmaped_intervals(ivs, len(seq))
    pieces = []
    for s, e in selected:
        if s < len(seq):
            pieces.append(seq[s:e])
        else:
            raise ValueError(f"InvalidInterval {s}..{e}")
    return pieces, selected


@dataclass
class Predict:
    """src/smooth/predict.rs:33-46."""
    prediction: List[int]
    seq: str
    id: str
    is_truncated: bool = False
    qual: Optional[str] = None

    def prediction_region(self):
        return get_label_region(self.prediction)

    def smooth_prediction(self, window_size):
        return get_label_region(majority_voting(self.prediction, window_size))

    def smooth_label(self, window_size):
        return majority_voting(self.prediction, window_size)

    def smooth_and_select_intervals(self, smooth_window_size, min_interval_size, approved_interval_number):
        return smooth_label_region(self.prediction, smooth_window_size, min_interval_size,
                                   approved_interval_number)


@dataclass
class FastqRecord:
    name: str
    description: str
    seq: str
    qual: str

    def to_text(self) -> str:
        # noodles fastq writer: "@name[ description]\nSEQ\n+\nQUAL\n" (src/output/writefq.rs:405-426)
        head = self.name if not self.description else f"{self.name} {self.description}"
        return f"@{head}\n{self.seq}\n+\n{self.qual}\n"


CHOP_TERMINAL, CHOP_INTERNAL, CHOP_ALL = "terminal", "internal", "all"


def _split_records_by_remove_internal(seq: str, rid: str, qual: str, target: Sequence[Interval],
                                      min_retain_interval_length: Optional[int]):
    """src/output/split.rs:60-136."""
    seqs, selected = remove_intervals_and_keep_left(seq, target)
    quals, _ = remove_intervals_and_keep_left(qual, target)
    if len(seqs) != len(quals):
        raise ValueError("NotSameLengthForQualityAndSequence")
    for s, q in zip(seqs, quals):
        if len(s) != len(q):
            raise ValueError("NotSameLengthForQualityAndSequence")
    ids = [f"{rid}|{selected[x][0]}:{selected[x][1]}" for x in range(len(seqs))]
    before = len(seqs)
    if min_retain_interval_length is not None:
        keep = [(i, s, q) for i, s, q in zip(ids, seqs, quals) if len(s) >= min_retain_interval_length]
        return before, [k[0] for k in keep], [k[1] for k in keep], [k[2] for k in keep]
    return before, ids, seqs, quals


def split_noodle_records_by_intervals(seq, rid, qual, target):
    """src/output/split.rs:138-169 (``--ocq``: emit the adapter pieces)."""
    return [FastqRecord(f"{rid}|{s}:{e}", "", seq[s:e], qual[s:e]) for s, e in target]


def split_noodle_records_by_remove_intervals(seq, rid, qual, target, min_chop_read_length,
                                             id_annotation, chop_type):
    """src/output/split.rs:171-226."""
    before, ids, seqs, quals = _split_records_by_remove_internal(seq, rid, qual, target, min_chop_read_length)
    current = CHOP_TERMINAL if before == 1 else CHOP_INTERNAL
    if ((chop_type == CHOP_TERMINAL and current == CHOP_INTERNAL)
            or (chop_type == CHOP_INTERNAL and current == CHOP_TERMINAL)
            or (len(seqs) > 0 and len(seqs[0]) == len(seq))):
        return [FastqRecord(rid, "", seq, qual)]
    out = []
    for i, s, q in zip(ids, seqs, quals):
        name = f"{i}|{'T' if current == CHOP_TERMINAL else 'I'}" if id_annotation else i
        out.append(FastqRecord(name, "", s, q))
    return out


@dataclass
class ChopOptions:
    """src/bin/predict.rs:19-78 (clap defaults)."""
    smooth_window_size: int = 21
    min_interval_size: int = 13
    approved_interval_number: int = 20
    max_process_intervals: int = 4
    min_read_length_after_chop: int = 20
    output_chopped_seqs: bool = False
    chop_type: str = CHOP_ALL


def process_record(fq: FastqRecord, predict: Optional[Predict], opt: ChopOptions) -> Optional[List[FastqRecord]]:
    """src/bin/predict.rs:137-187 -- one FASTQ record through the gating rules; None == dropped."""
    if predict is None:
        return None
    if len(predict.seq) < MIN_READ_LEN:
        return [fq]
    ivs = predict.smooth_and_select_intervals(opt.smooth_window_size, opt.min_interval_size,
                                              opt.approved_interval_number)
    if len(ivs) > opt.max_process_intervals or len(ivs) == 0:
        return [fq]
    if len(predict.seq) != len(fq.qual):
        return [fq]
    if opt.output_chopped_seqs:
        return split_noodle_records_by_intervals(predict.seq, fq.name, fq.qual, ivs)
    return split_noodle_records_by_remove_intervals(predict.seq, fq.name, fq.qual, ivs,
                                                    opt.min_read_length_after_chop, True, opt.chop_type)


def chop_records(fastq: Sequence[FastqRecord], predicts: dict, opt: ChopOptions) -> List[FastqRecord]:
    """src/bin/predict.rs:130-192,275-339 -- FASTQ order, records without prediction dropped."""
    out: List[FastqRecord] = []
    for rec in fastq:
        r = process_record(rec, predicts.get(rec.name), opt)
        if r is not None:
            out.extend(r)
    return out


# ---- chop coordinates in the flat form the C-ABI returns (same gating, numbers only) -------------

ACTION_PASSTHROUGH, ACTION_CHOP_T, ACTION_CHOP_I, ACTION_ADAPTERS, ACTION_UNCHOPPED = 0, 1, 2, 3, 4


def chop_coordinates(labels: Sequence[int], qual_len: Optional[int], opt: ChopOptions):
    """Numeric twin of process_record: returns (action, adapter_intervals, kept_intervals).

    ``labels`` is the per-base prediction of one read (len == decoded seq len);
    ``qual_len`` the FASTQ quality length (None == same as len(labels)).
    action PASSTHROUGH -> emit the FASTQ record verbatim; UNCHOPPED -> emit ``@{id}`` + decoded seq + qual, uncut; CHOP_T/CHOP_I -> emit ``kept`` pieces
    named ``{id}|s:e|T`` / ``|I``; ADAPTERS (--ocq) -> emit ``adapter`` pieces ``{id}|s:e``.
    """
    n = len(labels)
    if n < MIN_READ_LEN:
        return ACTION_PASSTHROUGH, [], []
    ivs = smooth_label_region(labels, opt.smooth_window_size, opt.min_interval_size,
                              opt.approved_interval_number)
    if len(ivs) > opt.max_process_intervals or len(ivs) == 0:
        return ACTION_PASSTHROUGH, ivs, []
    if qual_len is not None and qual_len != n:
        return ACTION_PASSTHROUGH, ivs, []
    if opt.output_chopped_seqs:
        return ACTION_ADAPTERS, ivs, []
    selected = generate_unmaped_intervals(sorted(ivs), n)
    before = len(selected)
    kept = [(s, e) for s, e in selected if e - s >= opt.min_read_length_after_chop]
    current = CHOP_TERMINAL if before == 1 else CHOP_INTERNAL
    if ((opt.chop_type == CHOP_TERMINAL and current == CHOP_INTERNAL)
            or (opt.chop_type == CHOP_INTERNAL and current == CHOP_TERMINAL)
            or (len(kept) > 0 and kept[0][1] - kept[0][0] == n)):
        return ACTION_UNCHOPPED, ivs, []
    return (ACTION_CHOP_T if current == CHOP_TERMINAL else ACTION_CHOP_I), ivs, kept


# ---- prediction-batch decode (.pt wire format) -------------------------------------------------

def load_predicts_from_batch(prediction, target, seq, id_arr, ignore_label: int = IGNORE_LABEL) -> dict:
    """src/smooth/predict.rs:263-317 on already-unpickled arrays.

    prediction: float [B,L,2]; argmax(2) with ties -> index 0 (candle argmax, :275);
    positions with target == ignore_label are dropped (:287-289); id row = [len, truncated, ascii...].
    """
    import numpy as np
    prediction = np.asarray(prediction)
    target = np.asarray(target)
    seq = np.asarray(seq)
    id_arr = np.asarray(id_arr)
    labels = (prediction[..., 1] > prediction[..., 0]).astype(np.int8)
    out = {}
    for b in range(labels.shape[0]):
        keep = target[b] != ignore_label
        n_id = int(id_arr[b][0])
        rid = ascii_list2str(id_arr[b][2:2 + n_id])
        out[rid] = Predict(prediction=labels[b][keep].tolist(), seq=id_list2seq(seq[b][keep]), id=rid,
                           is_truncated=bool(id_arr[b][1] != 0))
    return out


def parse_fastq_text(text: str) -> List[FastqRecord]:
    """Minimal 4-line FASTQ reader (noodles semantics: name = up to first space, rest = description)."""
    lines = text.split("\n")
    recs = []
    i = 0
    while i + 3 < len(lines) + 1 and i < len(lines) and lines[i].startswith("@"):
        head = lines[i][1:]
        name, _, desc = head.partition(" ")
        recs.append(FastqRecord(name, desc, lines[i + 1], lines[i + 3]))
        i += 4
    return recs


# ---- StatResult / collect_statistics_for_predicts (src/smooth/stat.rs:16-308) -----------------------------------------

FLANK_SIZE_COUNT_PLOYA = 5   # src/smooth/stat.rs:16


def collect_statistics_for_predicts(predicts, smooth_window_size: int, min_interval_size: int,
                                    approved_interval_number: int, internal_threshold: float, ploya_threshold: int):
    """src/smooth/stat.rs:222-308, literally: reads shorter than MIN_READ_LEN are skipped; the rayon reduce merges the
    per-read results in input order.  Returns a plain dict with the fields of StatResult (stat.rs:19-41)."""
    import numpy as np
    res = dict(predicts_with_chop=[], smooth_predicts_with_chop=[], smooth_internal_predicts=[], smooth_intervals={},
               original_intervals={}, total_truncated=0, smooth_only_one=[], smooth_only_one_with_ploya=[],
               total_predicts=0, smooth_intervals_relative_pos=[])
    for p in predicts:
        if len(p.seq) < MIN_READ_LEN:
            continue
        res["total_predicts"] += 1
        if p.is_truncated:
            res["total_truncated"] += 1
        regions = p.prediction_region()
        if regions:
            res["predicts_with_chop"].append(p.id)
            res["original_intervals"][p.id] = list(regions)
        smooth = p.smooth_and_select_intervals(smooth_window_size, min_interval_size, approved_interval_number)
        if smooth:
            res["smooth_predicts_with_chop"].append(p.id)
            res["smooth_intervals"][p.id] = list(smooth)
            if len(smooth) == 1:
                res["smooth_only_one"].append(p.id)
                s0 = smooth[0][0]
                count = p.seq[max(0, s0 - FLANK_SIZE_COUNT_PLOYA):s0].count("A")
                if count >= ploya_threshold:
                    res["smooth_only_one_with_ploya"].append(p.id)
            for (_, e) in smooth:
                rel = np.float32(e) / np.float32(len(p.seq))          # `region.1 as f32 / predict.seq_len() as f32`
                res["smooth_intervals_relative_pos"].append(float(rel))
                if rel < np.float32(internal_threshold):
                    res["smooth_internal_predicts"].append(p.id)
    return res
