"""K1 variant timing on BASELINE configs[1] shapes (28 x 20 min: 203,420 frames)."""
import json, sys
import numpy as np
sys.path.insert(0, ".")
from needle_b200 import engine, synth

ctx = engine.Context(0)
rng = np.random.default_rng(0)
segs = []
for e in range(28):
    n = 20 * 60 * 11025
    x = rng.integers(-8000, 8000, n, dtype=np.int16)
    segs += [x[: n // 2], x[3 * n // 4:]]
ps = engine.PcmSet.upload(ctx, segs)
frames = sum(synth.num_frames(s.size) for s in segs)
out = {}
ref = None
for variant in [int(v) for v in (sys.argv[1:] or ["1", "8", "10", "12"])]:
    ctx.set_option(2, variant)
    for _ in range(3):
        hs = ps.fingerprint()
    best = None
    for _ in range(7):
        hs = ps.fingerprint()
        ms = ctx.last_kernel_ms()["fp_fft_chroma"]
        best = ms if best is None else min(best, ms)
    h, t, off = hs.download()
    if ref is None:
        ref = h
    out["k1v%d" % variant] = dict(k1_ms=best, Mframes_per_s=frames / best / 1e3,
                                  fp32_TFLOPs=frames * 134.6e3 / (best * 1e-3) / 1e12,
                                  frac_of_74p4=frames * 134.6e3 / (best * 1e-3) / 74.44992e12,
                                  hashes_equal_to_v0=float(np.mean(h == ref)))
ctx.set_option(2, 0)
print(json.dumps(out, indent=1))
