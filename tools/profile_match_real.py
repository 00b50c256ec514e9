"""ncu target: the match kernels on REAL fingerprints of the bench's synthetic season
(BASELINE configs[1]: 28 x 20 min), not on uniform random hashes -- stationary background
audio makes unrelated frames match with p ~ 0.15-0.2 and neighbouring cells of a diagonal
correlate, which is what the adaptive kernel's later stages are for."""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
from needle_b200 import engine, synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = engine.Context(0)
made = bench.make_segments(list(range(28)), 28, 20.0)
segs, seeks = [], []
for v in range(28):
    a, b, sk = made[v]
    segs += [a, b]
    seeks += [0, sk]
ps = engine.PcmSet.upload(ctx, segs)
hs = ps.fingerprint(stride=2, seek_to_ns=seeks)
p = engine.match_params(include_endings=True)
for _ in range(iters):
    rs = hs.match(p)
    print(rs.count(), ctx.last_kernel_ms())
    rs.free()
