"""Seeded synthetic workloads for the configs in BASELINE.json (there is no
network and no decodable media here, so every test and benchmark runs on these).

Two levels:
  * PCM seasons  -- 11025 Hz i16 episodes: distinct low-passed noise with a
    shared "intro" and "credits" (chord sequences) spliced in at per-episode
    offsets.  Input of the fingerprint stage (what needle's Analyzer feeds to
    Chromaprint after FFmpeg decode + swresample, analyzer.rs:179-187,275).
  * hash seasons -- u32 sub-fingerprint lists with a planted shared run, each
    copy with random bit flips and occasional hard breaks.  Input of the match
    stage (what Comparator::run_with_frame_hashes receives, comparator.rs:524).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SAMPLE_RATE = 11025
FRAME = 4096
HOP = 1365
DELAY_NS = 2_600_000_000      # chromaprint_get_delay_ms = 28666*1000/11025 = 2600
ITEM_NS = 123_000_000         # chromaprint_get_item_duration_ms = 1365*1000/11025 = 123
HASH_DURATION_NS = 300_000_012  # Duration::from_secs_f32(0.3)


# ------------------------------------------------------------------- timing

def _f32(x):
    return np.float32(x)


def duration_from_secs_f32(x) -> np.ndarray:
    """Rust Duration::from_secs_f32 on an array of f32 (exact product in f64:
    24-bit significand x 5^9 fits 53 bits; rint = ties-to-even)."""
    return np.rint(np.asarray(x, dtype=np.float32).astype(np.float64) * 1e9).astype(np.uint64)


def duration_as_secs_f32(ns) -> np.ndarray:
    ns = np.asarray(ns, dtype=np.uint64)
    secs = (ns // np.uint64(1_000_000_000)).astype(np.float32)
    nanos = (ns % np.uint64(1_000_000_000)).astype(np.float32)
    return (secs + nanos / np.float32(1e9)).astype(np.float32)


def duration_mul_f32(ns: int, rhs) -> np.ndarray:
    s = duration_as_secs_f32(np.uint64(ns))
    return duration_from_secs_f32(np.asarray(rhs, dtype=np.float32) * s)


def hash_timestamps(n_raw: int, step_by: int = 2, delay_ns: int = DELAY_NS, item_ns: int = ITEM_NS,
                    seek_to_ns: int = 0) -> np.ndarray:
    """ts of the kept raw indices 0, step, 2*step, ... (analyzer.rs:293-318)."""
    idx = np.arange(0, n_raw, step_by, dtype=np.int64)
    return (np.uint64(delay_ns) + duration_mul_f32(item_ns, idx.astype(np.float32))
            + np.uint64(seek_to_ns)).astype(np.uint64)


def num_frames(n_samples: int) -> int:
    return (n_samples - FRAME) // HOP + 1 if n_samples >= FRAME else 0


def num_raw_hashes(n_samples: int) -> int:
    return max(num_frames(n_samples) - 19, 0)


# ---------------------------------------------------------------- PCM level

def _chords(rng: np.random.Generator, seconds: float, amp: float = 0.3) -> np.ndarray:
    """Sum of 6-12 random-pitch harmonic tones, new chord every 0.5-2 s."""
    n = int(round(seconds * SAMPLE_RATE))
    out = np.zeros(n, dtype=np.float64)
    pos = 0
    while pos < n:
        ln = min(int(rng.uniform(0.5, 2.0) * SAMPLE_RATE), n - pos)
        t = np.arange(ln) / SAMPLE_RATE
        k = int(rng.integers(6, 13))
        seg = np.zeros(ln)
        for _ in range(k):
            f0 = 55.0 * 2.0 ** (rng.integers(0, 48) / 12.0)      # A1 .. ~A5
            for h in (1, 2, 3):
                if f0 * h < 3400:
                    seg += np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 2 * np.pi)) / (h * k)
        out[pos:pos + ln] = seg
        pos += ln
    return out * amp


def _noise(rng: np.random.Generator, n: int, amp: float = 0.25, a: float = 0.9) -> np.ndarray:
    """1-pole low-passed white noise so the chroma vector is not flat."""
    from scipy.signal import lfilter
    x = rng.standard_normal(n)
    y = lfilter([1.0 - a], [1.0, -a], x)
    return y * (amp / max(y.std(), 1e-9))


@dataclass
class PcmEpisode:
    pcm: np.ndarray          # mono i16, whole episode
    intro_at: float          # seconds
    credits_at: float


def season_themes(season_seed: int, intro_s: float = 90.0, credits_s: float = 90.0):
    """The audio every episode of a season shares: (intro, credits) as float arrays."""
    srng = np.random.default_rng(season_seed)
    return _chords(srng, intro_s), _chords(srng, credits_s)


def make_pcm_episode(season_seed: int, e: int, minutes: float, intro: np.ndarray, credits: np.ndarray) -> PcmEpisode:
    """Episode e of a season: its own low-passed noise with the season's intro
    and credits spliced in at episode-specific offsets, +-1 LSB dither."""
    D = minutes * 60.0
    n = int(round(D * SAMPLE_RATE))
    intro_s, credits_s = intro.size / SAMPLE_RATE, credits.size / SAMPLE_RATE
    rng = np.random.default_rng(1000 * season_seed + e)
    x = _noise(rng, n)
    intro_at = float(rng.uniform(10.0, min(120.0, max(11.0, 0.5 * D - intro_s - 5.0))))
    hi = max(6.0, min(60.0, 0.25 * D - credits_s - 1.0))
    credits_at = float(D - credits_s - rng.uniform(5.0, hi))
    a = int(intro_at * SAMPLE_RATE)
    x[a:a + intro.size] = intro[:max(0, min(intro.size, n - a))]
    b = int(credits_at * SAMPLE_RATE)
    x[b:b + credits.size] = credits[:max(0, min(credits.size, n - b))]
    x = x * 32767.0 + rng.integers(-1, 2, n)     # per-episode +-1 LSB dither
    return PcmEpisode(np.clip(np.rint(x), -32768, 32767).astype(np.int16), intro_at, credits_at)


def make_pcm_season(n_episodes: int, minutes: float, season_seed: int = 1, intro_s: float = 90.0,
                    credits_s: float = 90.0) -> list[PcmEpisode]:
    intro, credits = season_themes(season_seed, intro_s, credits_s)
    return [make_pcm_episode(season_seed, e, minutes, intro, credits) for e in range(n_episodes)]


def split_segments(pcm: np.ndarray, opening_pct: float = 0.5, ending_pct: float = 0.25):
    """What Analyzer::run_single hashes (analyzer.rs:378-402): the first
    opening_pct of the stream, and from (1 - ending_pct) to the end.  Returns
    (opening_pcm, ending_pcm, ending_seek_to_ns); the seek uses Duration::mul_f32."""
    n = pcm.shape[0]
    dur_ns = int(round(n / SAMPLE_RATE * 1e9))
    open_ns = int(duration_mul_f32(dur_ns, np.float32(opening_pct)))
    seek_ns = int(duration_mul_f32(dur_ns, np.float32(1.0) - np.float32(ending_pct)))
    n_open = min(n, int(open_ns * SAMPLE_RATE // 1_000_000_000))
    s_end = min(n, int(seek_ns * SAMPLE_RATE // 1_000_000_000))
    return pcm[:n_open], pcm[s_end:], seek_ns


# --------------------------------------------------------------- hash level

@dataclass
class HashSeason:
    hashes: np.ndarray            # u32 concatenated: opening_0, ending_0, opening_1, ...
    ts_ns: np.ndarray             # u64
    seg_offset: np.ndarray        # u64 [2N+1]
    hash_duration_ns: np.ndarray  # u64 [N]

    @property
    def n_videos(self) -> int:
        return (self.seg_offset.size - 1) // 2

    def n_cells(self, include_endings: bool) -> int:
        """Algorithmic work of the match stage: one Hamming test per (i, j)
        of every pair's opening x opening (+ ending x ending) table."""
        off = self.seg_offset.astype(np.int64)
        no = off[1::2] - off[0:-1:2]
        ne = off[2::2] - off[1::2]
        tot = (no.sum() ** 2 - (no ** 2).sum()) // 2
        if include_endings:
            tot += (ne.sum() ** 2 - (ne ** 2).sum()) // 2
        return int(tot)


def _flip_bits(rng: np.random.Generator, h: np.ndarray, p: float) -> np.ndarray:
    mask = np.zeros(h.size, dtype=np.uint32)
    for b in range(32):
        mask |= (rng.random(h.size) < p).astype(np.uint32) << np.uint32(b)
    return h ^ mask


def make_hash_season(n_videos: int, n_open: int, n_end: int, seed: int = 0, run_len: int = 366,
                     flip_p: float = 0.08, break_p: float = 0.01, correlated: bool = False,
                     jitter_len: bool = False) -> HashSeason:
    """Uniform random u32 hashes with one shared run planted at a random offset
    in every opening and every ending list; each copy gets Binomial(32, flip_p)
    bit flips per hash and hard breaks with probability break_p.
    correlated=True makes the background temporally correlated (each hash =
    previous with 3 random flips): many more short accidental runs."""
    rng = np.random.default_rng(seed)
    shared_open = rng.integers(0, 2 ** 32, run_len, dtype=np.uint64).astype(np.uint32)
    shared_end = rng.integers(0, 2 ** 32, run_len, dtype=np.uint64).astype(np.uint32)
    hs, ts, off = [], [], [0]
    for v in range(n_videos):
        for (n, shared, seek) in ((n_open, shared_open, 0), (n_end, shared_end, 1_080_000_000_000)):
            if jitter_len and n > 8:
                n = int(n - rng.integers(0, min(64, n // 4)))
            if correlated:
                h = np.empty(n, dtype=np.uint32)
                cur = np.uint32(rng.integers(0, 2 ** 32))
                flips = rng.integers(0, 32, (n, 3))
                for k in range(n):
                    for b in flips[k]:
                        cur ^= np.uint32(1) << np.uint32(b)
                    h[k] = cur
            else:
                h = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
            if n > 2:
                L = min(run_len, n - 2)
                at = int(rng.integers(1, n - L + 1))
                copy = _flip_bits(rng, shared[:L].copy(), flip_p)
                brk = rng.random(L) < break_p
                copy[brk] = rng.integers(0, 2 ** 32, int(brk.sum()), dtype=np.uint64).astype(np.uint32)
                h[at:at + L] = copy
            hs.append(h)
            ts.append(hash_timestamps(2 * n, 2, seek_to_ns=seek if n else 0)[:n])
            off.append(off[-1] + n)
    return HashSeason(np.concatenate(hs), np.concatenate(ts), np.asarray(off, dtype=np.uint64),
                      np.full(n_videos, HASH_DURATION_NS, dtype=np.uint64))


SILENCE_HASH = 627964279   # Chromaprint TEST2 sub-fingerprint of digital silence (upstream tests/test_api.cpp)


def make_adversarial_season(kind: str, n_videos: int, n_open: int, n_end: int, seed: int = 7) -> HashSeason:
    """Seasons built to keep diagonals of the match matrix alive (what real TV audio does and uniform
    random hashes do not): `random` (the planted-run season), `correlated` (each hash = the previous
    with 3 bit flips), `silence60` (60 s of one constant hash somewhere in every list: a block of
    matching cells in every table), `jingle20` (a 40-hash jingle repeated 20 times in every list:
    thousands of runs below the 20 s minimum), `quiet_half` (the second half of every list silent),
    `all_silence` (every hash identical: every cell with i, j >= 1 matches)."""
    rng = np.random.default_rng(seed)
    if kind == "random":
        return make_hash_season(n_videos, n_open, n_end, seed=seed)
    if kind == "correlated":
        return make_hash_season(n_videos, n_open, n_end, seed=seed, correlated=True)
    s = make_hash_season(n_videos, n_open, n_end, seed=seed)
    off = s.seg_offset.astype(np.int64)
    jingle = rng.integers(0, 2 ** 32, 40, dtype=np.uint64).astype(np.uint32)
    for k in range(2 * n_videos):
        a, b = int(off[k]), int(off[k + 1])
        ln = b - a
        if ln < 4:
            continue
        if kind == "silence60":
            w = min(244, ln - 2)
            at = int(rng.integers(1, ln - w + 1))
            s.hashes[a + at:a + at + w] = SILENCE_HASH
        elif kind == "jingle20":
            if ln > 48:
                for _ in range(20):
                    at = int(rng.integers(1, ln - 41))
                    s.hashes[a + at:a + at + 40] = jingle
        elif kind == "quiet_half":
            s.hashes[a + ln // 2:b] = SILENCE_HASH
        elif kind == "all_silence":
            s.hashes[a:b] = SILENCE_HASH
        else:
            raise ValueError(kind)
    return s
