"""Where the host-buffer search (nb200_search) spends its wall time: upload, search of the resident
season, the whole call.  BASELINE configs[3] shapes; prints one JSON object."""
import json
import os
import sys
import time
sys.path.insert(0, ".")
import numpy as np
from needle_b200 import engine, synth

season = synth.make_hash_season(200, 2897, 1443, seed=4)
params = engine.match_params(include_endings=True)
ctx = engine.Context(0)
hd = season.hash_duration_ns


def wall(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def upload_only():
    hs = engine.HashSet.upload(ctx, season.hashes, season.ts_ns, season.seg_offset)
    ctx.synchronize()
    hs.free()


def upload_nosync():
    hs = engine.HashSet.upload(ctx, season.hashes, season.ts_ns, season.seg_offset)
    hs.free()


resident = engine.HashSet.upload(ctx, season.hashes, season.ts_ns, season.seg_offset)
out = {"cpus": os.cpu_count(), "bytes": int(season.hashes.nbytes + season.ts_ns.nbytes),
       "dtypes": [str(season.hashes.dtype), str(season.ts_ns.dtype), str(season.seg_offset.dtype)],
       "upload_sync_ms": wall(upload_only), "upload_return_ms": wall(upload_nosync),
       "search_resident_ms": wall(lambda: resident.search(hd, params)),
       "search_host_buffers_ms": wall(lambda: ctx.search(season.hashes, season.ts_ns, season.seg_offset, hd, params))}
ctx.host_profile(reset=True)
for _ in range(10):
    ctx.search(season.hashes, season.ts_ns, season.seg_offset, hd, params)
out["host_phases_ms_per_call"] = {k: v / 10 for k, v in ctx.host_profile().items()}
a = np.empty(season.ts_ns.size, np.uint64)
t0 = time.perf_counter()
for _ in range(20):
    np.copyto(a, season.ts_ns)
out["numpy_copy_GBs"] = 20 * a.nbytes / (time.perf_counter() - t0) / 1e9
print(json.dumps(out, indent=1))
