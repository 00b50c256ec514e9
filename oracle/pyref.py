"""Second, independent restatement of the reference path in plain Python/numpy.

TEST INFRASTRUCTURE ONLY (see oracle/needle_oracle.h).  It exists so that the C
oracle (match_ref.c, chromaprint_ref.c) is not the only transcription of
needle/src/audio/comparator.rs and of Chromaprint TEST2: tests/test_oracle_*.py
require the two to agree exactly on seeded inputs.  Pure-Python loops -- small
cases only.  Parity unpinned: the reference has no golden vectors.
"""
from __future__ import annotations

import math
import struct

import numpy as np

# ------------------------------------------------------------------ Duration


def f32(x: float) -> float:
    return struct.unpack("f", struct.pack("f", x))[0]


def duration_from_secs_f32(x: float) -> int:
    """Duration::from_secs_f32: exact, round to nearest ns, ties to even."""
    from fractions import Fraction
    fr = Fraction(f32(x)) * 1_000_000_000
    fl = fr.numerator // fr.denominator
    rem = fr - fl
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and fl % 2 == 1):
        fl += 1
    return fl


def duration_as_secs_f32(ns: int) -> float:
    secs, nanos = divmod(ns, 1_000_000_000)
    return f32(f32(float(secs)) + f32(f32(float(nanos)) / f32(1e9)))


def duration_mul_f32(ns: int, rhs: float) -> int:
    return duration_from_secs_f32(f32(f32(rhs) * duration_as_secs_f32(ns)))


def hash_timestamp(delay_ns: int, item_ns: int, i: int, seek_to_ns: int = 0) -> int:
    """analyzer.rs:309,314-318"""
    return delay_ns + duration_mul_f32(item_ns, f32(float(i))) + seek_to_ns


# --------------------------------------------------------------------- match


def simhash32(hashes) -> int:
    v = [0] * 32
    for h in hashes:
        h = int(h)
        for b in range(32):
            v[b] += 1 if (h >> b) & 1 else -1
    out = 0
    for b in range(32):
        if v[b] > 0:
            out |= 1 << b
    return out


def _heap_push(data: list, e: tuple):
    """BinaryHeap::push: append, sift up while element > parent (tuple order =
    derived lexicographic Ord of ComparatorHeapEntry, comparator.rs:20-35)."""
    data.append(e)
    pos = len(data) - 1
    while pos > 0:
        parent = (pos - 1) // 2
        if e[:13] <= data[parent][:13]:
            break
        data[pos] = data[parent]
        pos = parent
    data[pos] = e


def longest_common_hash_match(src, dst, threshold, min_opening_ns, min_ending_ns,
                              src_hash_duration_ns, dst_hash_duration_ns, is_opening):
    """comparator.rs:157-250.  src/dst: lists of (hash, ts_ns).  Returns the
    heap array; entries are tuples in the Rust field order followed by
    (i_end, j_end), which cannot change the order because (i, j) determines all
    preceding fields' tie-breaks only after every Rust field compared equal."""
    n, m = len(src), len(dst)
    if n == 0 or m == 0:
        return []
    is_ending = not is_opening
    heap: list = []
    table = [[0] * (m + 1) for _ in range(n + 1)]
    for i in range(n):
        for j in range(m):
            if i == 0 or j == 0:
                table[i][j] = 0
            elif bin(int(src[i][0]) ^ int(dst[j][0])).count("1") <= threshold:
                table[i][j] = table[i - 1][j - 1] + 1
            else:
                table[i][j] = 0
    for i in range(n - 1, 0, -1):
        for j in range(m - 1, 0, -1):
            if table[i][j] == 0 or (i < n - 1 and j < m - 1 and table[i + 1][j + 1] != 0):
                continue
            ln = table[i][j]
            ss, se = i - ln, i
            ds, de = j - ln, j
            src_start, src_end = int(src[ss][1]), int(src[se][1])
            dst_start, dst_end = int(dst[ds][1]), int(dst[de][1])
            if src_end < src_start or dst_end < dst_start:
                raise OverflowError("Duration underflow")
            mn = min_opening_ns if is_opening else min_ending_ns
            if not ((src_end - src_start) >= mn and (dst_end - dst_start) >= mn):
                continue
            smh = simhash32([h for h, _ in src[ss:se + 1]])
            dmh = simhash32([h for h, _ in dst[ds:de + 1]])
            e = (ln, src_start, src_end, dst_start, dst_end, smh, dmh,
                 is_opening, is_ending, is_opening, is_ending,
                 src_hash_duration_ns, dst_hash_duration_ns)
            _heap_push(heap, e + (i, j))
    return heap


def run_with_frame_hashes(videos, threshold=10, min_opening_ns=20_000_000_000,
                          min_ending_ns=20_000_000_000, time_padding_ns=0, include_endings=False):
    """comparator.rs:524-629.  videos: list of dicts {opening: [(h, ts)],
    ending: [(h, ts)], hash_duration_ns}.  Returns per-video result tuples
    (present, has_opening, has_ending, o_start, o_end, e_start, e_end)."""
    N = len(videos)
    pairs = []
    processed = [False] * N
    for i in range(N):
        for j in range(N):
            if i == j or processed[j]:
                continue
            pairs.append((i, j))
        processed[i] = True

    data = []
    for (s, d) in pairs:
        vs, vd = videos[s], videos[d]
        op = longest_common_hash_match(vs["opening"], vd["opening"], threshold, min_opening_ns,
                                       min_ending_ns, vs["hash_duration_ns"], vd["hash_duration_ns"], True)
        en = []
        if include_endings:
            if len(vs["ending"]) == 0 or len(vd["ending"]) == 0:
                raise ValueError("FrameHashDataNoEnding")
            en = longest_common_hash_match(vs["ending"], vd["ending"], threshold, min_opening_ns,
                                           min_ending_ns, vs["hash_duration_ns"], vd["hash_duration_ns"], False)
        if op or en:
            data.append((s, d, op, en))

    info_map = [[] for _ in range(N)]
    for (s, d, op, en) in data:
        info_map[s].append((op, en, True))
        info_map[d].append((op, en, False))

    results = []
    for v in range(N):
        matches = info_map[v]
        if not matches:
            results.append((0, 0, 0, 0, 0, 0, 0))
            continue
        cands = []
        for (op, en, is_source) in matches:
            for lst, is_op in ((op, True), (en, False)):
                for e in lst:
                    if is_source:
                        cands.append(((e[1], e[2]), e[11], e[5], is_op))
                    else:
                        cands.append(((e[3], e[4]), e[12], e[6], is_op))
        distinct = {}
        bias = threshold + threshold // 2
        for i, c in enumerate(cands):
            for j, o in enumerate(cands):
                if bin(c[2] ^ o[2]).count("1") >= bias:
                    continue
                distinct.setdefault(i, set()).add(j)
                distinct.setdefault(j, set()).add(i)
        res = [1, 0, 0, 0, 0, 0, 0]
        for want_opening in (True, False):
            if not want_opening and not include_endings:
                break
            best = []
            for k, vset in distinct.items():
                if cands[k][3] != want_opening:
                    continue
                (start, end) = cands[k][0]
                count = len(vset)
                dur = duration_as_secs_f32(end - start)
                score = f32(-f32(f32(f32(float(count)) * f32(0.3)) + f32(dur * f32(0.7))))
                best.append((score, k))
            best.sort()
            if best:
                k = best[0][1]
                (start, end), hd = cands[k][0], cands[k][1]
                if want_opening:
                    res[1], res[3], res[4] = 1, start + time_padding_ns, end - time_padding_ns - hd
                else:
                    res[2], res[5], res[6] = 1, start + time_padding_ns, end - time_padding_ns - hd
        results.append(tuple(res))
    return results


# --------------------------------------------------------------- fingerprint

FRAME = 4096
HOP = 1365
CLASSIFIERS = [
    (0, 4, 3, 15, 1.98215, 2.35817, 2.63523),
    (4, 4, 6, 15, -1.03809, -0.651211, -0.282167),
    (1, 0, 4, 16, -0.298702, 0.119262, 0.558497),
    (3, 8, 2, 12, -0.105439, 0.0153946, 0.135898),
    (3, 4, 4, 8, -0.142891, 0.0258736, 0.200632),
    (4, 0, 3, 5, -0.826319, -0.590612, -0.368214),
    (1, 2, 2, 9, -0.557409, -0.233035, 0.0534525),
    (2, 7, 3, 4, -0.0646826, 0.00620476, 0.0784847),
    (2, 6, 2, 16, -0.192387, -0.029699, 0.215855),
    (2, 1, 3, 2, -0.0397818, -0.00568076, 0.0292026),
    (5, 10, 1, 15, -0.53823, -0.369934, -0.190235),
    (3, 6, 2, 10, -0.124877, 0.0296483, 0.139239),
    (2, 1, 1, 14, -0.101475, 0.0225617, 0.231971),
    (3, 5, 6, 4, -0.0799915, -0.00729616, 0.063262),
    (1, 9, 2, 12, -0.272556, 0.019424, 0.302559),
    (3, 4, 2, 14, -0.164292, -0.0321188, 0.0846339),
]
GRAY = [0, 1, 3, 2]


def chroma_notes():
    lo = max(1, int(round(FRAME * 28 / 11025)))
    hi = min(FRAME // 2, int(round(FRAME * 3520 / 11025)))
    notes = np.zeros(FRAME, dtype=np.int64)
    for i in range(lo, hi):
        freq = i * 11025 / FRAME
        octave = math.log(freq / (440.0 / 16.0)) / math.log(2.0)
        notes[i] = int(12 * (octave - math.floor(octave)))
    return lo, hi, notes


def fingerprint(pcm: np.ndarray, channels: int = 1):
    """Chromaprint TEST2 raw sub-fingerprints; the FFT is numpy's (pocketfft),
    i.e. independent of chromaprint_ref.c's Stockham FFT.  Window sums are
    direct (no integral image): mathematically the same areas."""
    x = np.asarray(pcm, dtype=np.int16).reshape(-1)
    if channels == 2:
        lr = x.reshape(-1, 2).astype(np.int32)
        s = lr[:, 0] + lr[:, 1]
        x = (np.sign(s) * (np.abs(s) // 2)).astype(np.int16)  # C division truncates toward zero
    n = x.size
    nf = (n - FRAME) // HOP + 1 if n >= FRAME else 0
    win = (1.0 / 32767.0) * (0.54 - 0.46 * np.cos(np.arange(FRAME) * 2.0 * np.pi / (FRAME - 1)))
    lo, hi, notes = chroma_notes()
    chroma = np.zeros((nf, 12))
    for f in range(nf):
        spec = np.fft.rfft(x[f * HOP:f * HOP + FRAME].astype(np.float64) * win)
        p = spec.real ** 2 + spec.imag ** 2
        chroma[f] = np.bincount(notes[lo:hi], weights=p[lo:hi], minlength=12)
    if nf < 20:
        return np.zeros(0, np.uint32), chroma
    co = [0.25, 0.75, 1.0, 0.75, 0.25]
    rows = sum(co[j] * chroma[j:nf - 4 + j] for j in range(5))
    norm = np.sqrt((rows ** 2).sum(axis=1))
    rows = np.where(norm[:, None] < 0.01, 0.0, rows / np.where(norm[:, None] == 0, 1, norm[:, None]))
    out = np.zeros(nf - 19, np.uint32)

    def area(x0, r1, c1, r2, c2):
        return rows[x0 + r1:x0 + r2, c1:c2].sum()

    for x0 in range(nf - 19):
        bits = 0
        for (t, y, h, w, t0, t1, t2) in CLASSIFIERS:
            if t == 0:
                a, b = area(x0, 0, y, w, y + h), 0.0
            elif t == 1:
                h2 = h // 2
                a, b = area(x0, 0, y + h2, w, y + h), area(x0, 0, y, w, y + h2)
            elif t == 2:
                w2 = w // 2
                a, b = area(x0, w2, y, w, y + h), area(x0, 0, y, w2, y + h)
            elif t == 3:
                w2, h2 = w // 2, h // 2
                a = area(x0, 0, y + h2, w2, y + h) + area(x0, w2, y, w, y + h2)
                b = area(x0, 0, y, w2, y + h2) + area(x0, w2, y + h2, w, y + h)
            elif t == 4:
                h3 = h // 3
                a = area(x0, 0, y + h3, w, y + 2 * h3)
                b = area(x0, 0, y, w, y + h3) + area(x0, 0, y + 2 * h3, w, y + h)
            else:
                w3 = w // 3
                a = area(x0, w3, y, 2 * w3, y + h)
                b = area(x0, 0, y, w3, y + h) + area(x0, 2 * w3, y, w, y + h)
            v = math.log((1.0 + a) / (1.0 + b))
            q = (0 if v < t0 else 1) if v < t1 else (2 if v < t2 else 3)
            bits = ((bits << 2) | GRAY[q]) & 0xFFFFFFFF
        out[x0] = bits
    return out, chroma
