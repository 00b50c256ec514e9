"""ncu target for the match stage alone: a slice of BASELINE configs[3] (60 episodes of the search
leg's shapes, 1770 pairs with endings = 18.6 G cells), resident hashes.  argv[1] = iterations."""
import sys
sys.path.insert(0, ".")
from needle_b200 import engine, synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = engine.Context(0)
season = synth.make_hash_season(60, 2897, 1443, seed=4)
hs = engine.HashSet.upload(ctx, season.hashes, season.ts_ns, season.seg_offset)
p = engine.match_params(include_endings=True)
for _ in range(iters):
    rs = hs.match(p)
    print(rs.count(), ctx.last_kernel_ms())
    rs.free()
