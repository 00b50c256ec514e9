"""Where one search step goes when the match is short (one rank's 1/8 pair slice of configs[3], world of
one): wall time per nb200_mjob_run against the CUDA-event phases and the host-side phase timers."""
import json
import sys
import time
sys.path.insert(0, ".")
import numpy as np
from needle_b200 import engine, synth

season = synth.make_hash_season(200, 2897, 1443, seed=4)
params = engine.match_params(include_endings=True)
ctx = engine.Context(0)
comm = engine.Comm.init_rank(ctx, None, 0, 1)
seg_len = np.diff(season.seg_offset.astype(np.int64)).astype(np.uint64)
cuts = engine.plan_pairs(seg_len, 8, True)
allp = np.array([(a, b) for a in range(200) for b in range(a + 1, 200)], np.uint32)
out = {}
for name, pairs in (("slice_1_of_8", allp[int(cuts[0]):int(cuts[1])]), ("all_pairs", None)):
    job = engine.MultiJob.search([comm], season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, params,
                                 pairs=pairs)
    for _ in range(20):
        job.run()
    ctx.host_profile(reset=True)
    t0 = time.perf_counter()
    n = 100
    for _ in range(n):
        job.run()
    wall = (time.perf_counter() - t0) / n * 1e3
    host = {k: v / n for k, v in ctx.host_profile().items() if v}
    out[name] = {"wall_ms_per_run": wall, "gpu_phase_ms": job.phase_ms(), "kernel_ms": ctx.last_kernel_ms(),
                 "host_phase_ms_per_run": host}
    job.free()
print(json.dumps(out, indent=1))
