"""Match time of every rank's pair slice of BASELINE configs[3] at world = 8, measured one after the other
on ONE GPU (the kernel's tail and balance without an 8-GPU box).  Prints one JSON object."""
import json
import sys
sys.path.insert(0, ".")
import numpy as np
from needle_b200 import engine, synth

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
season = synth.make_hash_season(200, 2897, 1443, seed=4)
ctx = engine.Context(0)
hs = engine.HashSet.upload(ctx, season.hashes, season.ts_ns, season.seg_offset)
p = engine.match_params(include_endings=True)
seg_len = np.diff(season.seg_offset.astype(np.int64)).astype(np.uint64)
cuts = engine.plan_pairs(seg_len, world, True)
allp = np.array([(a, b) for a in range(200) for b in range(a + 1, 200)], np.uint32)
out = {"world": world, "slices_ms": []}
for r in range(world):
    pairs = allp[int(cuts[r]):int(cuts[r + 1])]
    best = 1e9
    for _ in range(6):
        rs = hs.match(p, pairs=pairs)
        best = min(best, ctx.last_kernel_ms()["match"])
        rs.free()
    out["slices_ms"].append(best)
best = 1e9
for _ in range(4):
    rs = hs.match(p)
    best = min(best, ctx.last_kernel_ms()["match"])
    rs.free()
out["whole_ms"] = best
out["ideal_slice_ms"] = best / world
out["max_slice_ms"] = max(out["slices_ms"])
print(json.dumps(out))
