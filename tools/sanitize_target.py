"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from needle_b200 import engine, synth

ctx = engine.Context(0)
rng = np.random.default_rng(0)
# fingerprint: ragged segments, every K1 variant, stereo, pipelined host path
lens = [0, 4095, 4096, 4096 + 1365 * 19, 4096 + 1365 * 40 + 7, 60_001, 90_000, 33_333]
segs = [rng.integers(-20000, 20000, n).astype(np.int16) for n in lens]
ref = None
for variant in (0, 1, 8, 10, 12, 16, 17, 18, 19, 112):
    ctx.set_option(2, variant)
    out = ctx.fingerprint_batch(segs)
    if ref is None:
        ref = out
    agree = sum(int(np.sum(a == b)) for a, b in zip(out, ref))
    total = sum(a.size for a in ref)
    assert agree >= 0.99 * total, (variant, agree, total)
ctx.set_option(2, 0)
st = np.stack([segs[5], segs[5]], axis=1).reshape(-1)
ctx.fingerprint_batch([st], channels=2)
p = engine.match_params(include_endings=True, min_opening_ns=3_000_000_000, min_ending_ns=2_000_000_000)
res = ctx.analyze_search([segs[6], segs[5], segs[6], segs[7], segs[5], segs[7]], 1, None, synth.HASH_DURATION_NS, p)
# match: general and fast kernels, small season with planted runs
season = synth.make_hash_season(4, 700, 300, seed=2, run_len=200)
for params in (engine.match_params(include_endings=True),
               engine.match_params(include_endings=True, min_opening_ns=0, min_ending_ns=10 ** 9, threshold=12)):
    runs = ctx.match_pairs(season.hashes, season.ts_ns, season.seg_offset, params)
    ctx.search(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, params)
# band groups of the fast kernel, page-locked input (direct copy + device-side move into the aligned layout)
long_season = synth.make_hash_season(3, 1100, 600, seed=5, run_len=300, jitter_len=True)
pin_h = engine.PinnedArray.empty(long_season.hashes.size, np.uint32)
pin_t = engine.PinnedArray.empty(long_season.ts_ns.size, np.uint64)
pin_h.array[:] = long_season.hashes
pin_t.array[:] = long_season.ts_ns
want = ctx.match_pairs(long_season.hashes, long_season.ts_ns, long_season.seg_offset, engine.match_params(include_endings=True))
for group in (2, 16):
    ctx.set_option(6, group)
    got = ctx.match_pairs(pin_h.array, pin_t.array, long_season.seg_offset, engine.match_params(include_endings=True))
    assert np.array_equal(got, want)
ctx.set_option(6, 0)
pin_h.free()
pin_t.free()
# device vote vs host vote, the PCM-resident fused call, run blocks (export -> vote_blocks)
hs = engine.HashSet.upload(ctx, season.hashes, season.ts_ns, season.seg_offset)
for params in (engine.match_params(include_endings=True),
               engine.match_params(include_endings=True, min_opening_ns=10 ** 9, min_ending_ns=10 ** 9, threshold=12)):
    dev = hs.search(season.hash_duration_ns, params)
    ctx.set_option(4, 1)
    host = hs.search(season.hash_duration_ns, params)
    ctx.set_option(4, 0)
    assert dev == host
import torch
block = 64 * (1 + 256)
buf = torch.zeros(2 * block, dtype=torch.uint8, device="cuda:0")
pairs = np.array([(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)], np.uint32)
params = engine.match_params(include_endings=True)
hs.match_export(params, pairs[:3], 0, buf.data_ptr(), block)
hs.match_export(params, pairs[3:], 3, buf.data_ptr() + block, block)
got, found, trunc = ctx.vote_blocks(buf.data_ptr(), 2, block, season.hash_duration_ns, params)
assert not trunc and got == hs.search(season.hash_duration_ns, params)
ps = engine.PcmSet.upload(ctx, [segs[6], segs[5], segs[6], segs[7], segs[5], segs[7]])
assert ps.search(None, synth.HASH_DURATION_NS, p) == res
print("sanitize target ok", len(runs), res[:1])
