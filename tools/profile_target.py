"""Small fixed workload for ncu captures: the BASELINE configs[1] shapes
(28 episodes x 20 min: 56 PCM segments, 378 pairs with endings), random PCM.
Launch order per iteration: fp_fft_chroma, fp_classify, match, simhash."""
import sys
import numpy as np
sys.path.insert(0, ".")
from needle_b200 import engine, synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = engine.Context(0)
if len(sys.argv) > 2:
    ctx.set_option(2, int(sys.argv[2]))     # NB200_OPT_K1_VARIANT
rng = np.random.default_rng(0)
n = 20 * 60 * 11025
segs = []
for e in range(28):
    x = rng.integers(-8000, 8000, n, dtype=np.int16)
    segs += [x[: n // 2], x[3 * n // 4:]]
ps = engine.PcmSet.upload(ctx, segs)
season = synth.make_hash_season(28, 2413, 1201, seed=1)
hs2 = engine.HashSet.upload(ctx, season.hashes, season.ts_ns, season.seg_offset)
p = engine.match_params(include_endings=True)
for _ in range(iters):
    hs = ps.fingerprint()
    hs.free()
    rs = hs2.match(p)
    print(rs.count(), ctx.last_kernel_ms())
    rs.free()
