"""Quick timings of the other BASELINE configs on one GPU (development aid, not the bench
contract; bench.py measures configs[3], [4] and [1]).  Prints one JSON object.

  config 3: 12 x 60 min, 66 pairs, opening + ending search (hashes resident)
  config 4: search-only, 200 x 24 min, 19,900 pairs (hashes resident; also from host arrays)
  config 5: fingerprint-only, N hours of 11025 Hz mono PCM resident in HBM
"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from needle_b200 import engine, synth

ctx = engine.Context(0)
out = {}
p = engine.match_params(include_endings=True)


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    ws, ks = [], []
    for _ in range(n):
        t0 = time.perf_counter()
        r = fn()
        ws.append((time.perf_counter() - t0) * 1e3)
        ks.append(ctx.last_kernel_ms())
    i = int(np.argmin(ws))
    return r, ws[i], ks[i]


for name, (n_videos, n_open, n_end) in {"config3_12x60min": (12, 7259, 3624),
                                        "config2_28x20min": (28, 2413, 1201),
                                        "config4_200x24min": (200, 2897, 1443)}.items():
    season = synth.make_hash_season(n_videos, n_open, n_end, seed=1)
    hs = engine.HashSet.upload(ctx, season.hashes, season.ts_ns, season.seg_offset)
    cells = season.n_cells(True)
    pairs = n_videos * (n_videos - 1) // 2
    res, wall, k = timed(lambda: hs.search(season.hash_duration_ns, p))
    _, wall_host, _ = timed(lambda: ctx.search(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, p))
    ctx.set_option(engine.Context.OPT_HOST_VOTE, 1)
    res_h, wall_hv, _ = timed(lambda: hs.search(season.hash_duration_ns, p))
    ctx.set_option(engine.Context.OPT_HOST_VOTE, 0)
    assert res == res_h
    out[name] = dict(pairs=pairs, cells=cells, openings=sum(r[1] for r in res), endings=sum(r[2] for r in res),
                     search_resident_ms=wall, search_from_host_arrays_ms=wall_host, search_host_vote_ms=wall_hv,
                     match_kernel_ms=k["match"], simhash_ms=k["simhash"], vote_kernels_ms=k["vote"],
                     pairs_per_s=pairs / wall * 1e3, Tcells_per_s_kernel=cells / k["match"] / 1e9)
    hs.free()

# config 5: fingerprint-only.  A pool of 25 distinct 24-minute episodes, tiled to `hours`.
hours = float(sys.argv[1]) if len(sys.argv) > 1 else 250.0
rng = np.random.default_rng(0)
pool = []
for e in range(25):
    n = 24 * 60 * 11025
    x = rng.integers(-8000, 8000, n, dtype=np.int16)
    pool.append((x[: n // 2], x[3 * n // 4:]))
per_ep_h = (pool[0][0].size + pool[0][1].size) / 11025 / 3600
n_eps = int(hours / per_ep_h)
segs = []
for e in range(n_eps):
    segs += list(pool[e % 25])
t0 = time.perf_counter()
ps = engine.PcmSet.upload(ctx, segs)
upload_s = time.perf_counter() - t0
frames = sum(synth.num_frames(s.size) for s in segs)
audio_h = sum(s.size for s in segs) / 11025 / 3600


def fp():
    h = ps.fingerprint()
    h.free()


_, wall, k = timed(fp, n=3, warm=1)
out["config5_fingerprint_only"] = dict(
    audio_hours=audio_h, episodes=n_eps, frames=frames, pcm_gb=sum(s.size for s in segs) * 2 / 1e9,
    upload_s_pageable=upload_s, wall_ms=wall, k1_ms=k["fp_fft_chroma"], k2_ms=k["fp_classify"],
    audio_hours_per_s=audio_h / wall * 1e3, Mframes_per_s=frames / k["fp_fft_chroma"] / 1e3,
    fp32_TFLOPs=frames * 134.6e3 / (k["fp_fft_chroma"] * 1e-3) / 1e12)
print(json.dumps(out, indent=1))
