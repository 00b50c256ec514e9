"""Quick kernel timing on the GPU box (development aid, not the bench contract)."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from needle_b200 import engine, synth

ctx = engine.Context(0)
out = {}
# match: N episodes x 24 min
for n_videos in (28, 60, 120):
    season = synth.make_hash_season(n_videos, 2897, 1443, seed=1)
    hs = engine.HashSet.upload(ctx, season.hashes, season.ts_ns, season.seg_offset)
    p = engine.match_params(include_endings=True)
    for name, opt in (("sampled", None), ("dense", 3), ("general", 1)):
        if opt:
            ctx.set_option(opt, 1)
        for _ in range(3):
            rs = hs.match(p)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); rs = hs.match(p); t1 = time.perf_counter()
            ms = ctx.last_kernel_ms()
            ts.append((ms["match"], ms["simhash"], (t1 - t0) * 1e3))
        if opt:
            ctx.set_option(opt, 0)
        n_runs, cells = rs.count()
        k = min(t[0] for t in ts)
        out["match_%d_%s" % (n_videos, name)] = dict(
            pairs=n_videos * (n_videos - 1) // 2, cells=cells, runs=n_runs, kernel_ms=k,
            simhash_ms=min(t[1] for t in ts), wall_ms=min(t[2] for t in ts), Tcells_per_s=cells / k / 1e9)
# fingerprint: 28 x 20 min (opening 50% + ending 25%)
rng = np.random.default_rng(0)
segs = []
for e in range(28):
    n = 20 * 60 * 11025
    x = rng.integers(-8000, 8000, n, dtype=np.int16)
    segs += [x[: n // 2], x[3 * n // 4:]]
t0 = time.perf_counter()
ps = engine.PcmSet.upload(ctx, segs)
t1 = time.perf_counter()
frames = sum(synth.num_frames(s.size) for s in segs)
for variant in (0, 4, 5, 6):
    ctx.set_option(2, variant)
    for _ in range(3):
        hs = ps.fingerprint()
    ts = []
    for _ in range(5):
        t2 = time.perf_counter(); hs = ps.fingerprint(); t3 = time.perf_counter()
        ms = ctx.last_kernel_ms()
        ts.append((ms["fp_fft_chroma"], ms["fp_classify"], (t3 - t2) * 1e3))
    k1 = min(t[0] for t in ts)
    out["fingerprint_28x20min_k1v%d" % variant] = dict(
        frames=frames, audio_hours=sum(s.size for s in segs) / 11025 / 3600,
        upload_ms=(t1 - t0) * 1e3, k1_ms=k1, k2_ms=min(t[1] for t in ts),
        wall_ms=min(t[2] for t in ts), Mframes_per_s=frames / k1 / 1e3,
        fp32_TFLOPs=frames * 134.6e3 / (k1 * 1e-3) / 1e12)
ctx.set_option(2, 0)
print(json.dumps(out, indent=1))
