#!/usr/bin/env python
"""bench.py -- the fingerprint-and-match path on the configurations BASELINE.json quotes its
metric on ("episode-pairs/sec (search) & audio-hours/sec fingerprinted, 1/2/4/8 B200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Three legs, every one a FIXED job split over the N GPUs (strong scaling), all through the
C ABI of include/needle_b200.h (multi-GPU: nb200_comm_* / nb200_mjob_*):

  search       BASELINE configs[3]: search-only from precomputed hashes, 200 episodes x 24 min,
               19,900 pairs, openings + endings; the pair list is sharded over the ranks by table
               cells, every rank's runs are pushed to rank 0 over NVLink, rank 0 votes.
               -> `value` (pairs/s, hashes resident in HBM) and `e2e` (page-locked host hash +
               timestamp arrays in, per-video results out, every step: nb200_search; at N > 1 a job
               created -- every rank copies 1/N of the season, one all-gather over NVLink -- run and
               freed per step); `e2e.pageable_input` = the same from ordinary memory.
  fingerprint  BASELINE configs[4]: fingerprint-only, 1000 audio-hours of 11025 Hz mono PCM
               (3,334 episodes x 24 min, opening 50 % + ending 25 % of each = 18 min), episodes
               sharded over the ranks, PCM resident in HBM -> `fingerprint.value`
               (audio-hours/s); `fingerprint.e2e` streams a bounded sample from pinned host memory.
  season       BASELINE configs[1]: 28 episodes x 20 min, analyze + search with endings from
               pinned host PCM, episodes and pairs sharded over the ranks -> `season_e2e`.

One JSON line on rank 0.  `roofline` is the binding roof of the kernel that takes most of
the GPU time of the whole run (K1 fp_fft_chroma: FP32, against the MEASURED FMA-pipe rate);
`roofline_popc` the match kernel's.  `--impl reference` times the CPU restatement of the
reference path (oracle/, all host threads) on bounded samples of the same jobs: the
reference itself is Rust and cannot be built in this image (no cargo).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from needle_b200 import synth  # noqa: E402

# ---- BASELINE configs[3]: search-only
SEARCH_EPISODES, SEARCH_N_OPEN, SEARCH_N_END = 200, 2897, 1443    # 24-min episodes: stored hashes per list
# ---- BASELINE configs[4]: fingerprint-only
FP_HOURS = 1000.0
FP_EPISODE_MIN = 24.0
FP_POOL = 24                  # distinct synthetic episodes; the job tiles them on the device
# ---- BASELINE configs[1]: analyze + search
SEASON_EPISODES, SEASON_MINUTES = 28, 20.0

FLOP_PER_FRAME = 134.6e3        # 4096-pt real FFT + window + power + fold + classify
BYTES_PER_FRAME = 1365 * 2 + 4  # mono i16 in (one hop) + one u32 hash out
# measured on this pool's B200s (profiles/r02_k1_fp_mix_peak.jsonl, tools/k1_mix_peak.cu): FP32 results per
# clock per SM that K1's own instruction mix reaches with nothing else in the way (nominal 128)
FP32_LANE_OPS_PER_CLK_PER_SM = 121.03
# measured POPC issue rate per clock per SM (profiles/r01_pipe_peak_warm.jsonl, tools/pipe_peak.cu; nominal 16)
POPC_PER_CLK_PER_SM = 15.0
SM_COUNT, SM_MHZ_MAX = 148, 1965.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--fp-hours", type=float, default=FP_HOURS, help="size of the fingerprint-only job (audio-hours)")
    ap.add_argument("--cpu-baseline", type=int, default=1, help="time the CPU oracle beside the GPU run (N=1)")
    ap.add_argument("--legs", default="search,fingerprint,season")
    return ap.parse_args()


# ------------------------------------------------------------------ workloads

def ncu_metric(path, name):
    """One number out of a committed ncu summary (tools/ncu_summary.py format); None if absent."""
    try:
        for ln in open(os.path.join(os.path.dirname(os.path.abspath(__file__)), path)):
            f = ln.split()
            if len(f) >= 2 and f[0] == name:
                return float(f[1])
    except (OSError, ValueError):
        pass
    return None


def search_season():
    """configs[3]: uniform random hashes, a shared run of 366 hashes (90 s) planted in every
    opening and ending list with bit flips and hard breaks."""
    return synth.make_hash_season(SEARCH_EPISODES, SEARCH_N_OPEN, SEARCH_N_END, seed=4)


def sub_season(season, n):
    off = season.seg_offset
    e = int(off[2 * n])
    return synth.HashSeason(season.hashes[:e].copy(), season.ts_ns[:e].copy(), off[:2 * n + 1].copy(),
                            season.hash_duration_ns[:n].copy())


def season_cells(season, include_endings=True):
    ln = np.diff(season.seg_offset.astype(np.int64))
    o, e = ln[0::2], ln[1::2]
    cells = (o.sum() ** 2 - (o ** 2).sum()) // 2
    if include_endings:
        cells += (e.sum() ** 2 - (e ** 2).sum()) // 2
    return int(cells)


def make_pcm_segments(video_ids, season_seed, minutes):
    """{video: (opening_pcm, ending_pcm, ending_seek_ns)} of one synthetic season."""
    themes = synth.season_themes(season_seed)

    def one(v):
        ep = synth.make_pcm_episode(season_seed, v, minutes, *themes)
        return v, synth.split_segments(ep.pcm)

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:
        return dict(ex.map(one, video_ids))


def segment_metadata(n_videos, minutes):
    n = int(round(minutes * 60.0 * synth.SAMPLE_RATE))
    a, b, seek = synth.split_segments(np.zeros(n, np.int16))
    return [a.size, b.size] * n_videos, [0, seek] * n_videos


# -------------------------------------------------------------------- clocks

class ClockSampler:
    """SM clock + throttle reasons sampled through NVML while `active`."""

    def __init__(self, index):
        self.samples, self.reasons, self.active, self._stop = [], set(), False, False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop:
            if self.active:
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
            time.sleep(0.005)

    def summary(self):
        self._stop = True
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------- reference arm

def cpu_search(orc, season, threads):
    s = orc.Season(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns)
    st, res, _ = orc.run_with_frame_hashes(s, include_endings=True, n_threads=threads)
    assert st == 0
    return [tuple(int(x) for x in r) for r in res]


def cpu_fingerprint(orc, segments, threads):
    return orc.fingerprint_many(segments, channels=1, n_threads=threads)


def run_reference(args):
    """The CPU restatement of the reference path on this box's host cores, on the b200 arm's
    config (BASELINE configs[3] search-only), each step a bounded sample of the 19,900 pairs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.lib()
    threads = os.cpu_count() or 1
    full = search_season()
    n_sub = 24
    t0 = time.perf_counter()
    cpu_search(orc, sub_season(full, n_sub), threads)
    probe = time.perf_counter() - t0
    per_pair = probe / (n_sub * (n_sub - 1) // 2)
    budget_s = 150.0
    n = SEARCH_EPISODES
    while n > 8 and per_pair * (n * (n - 1) // 2) * (args.steps + args.warmup) > budget_s:
        n -= 1
    season = sub_season(full, n)
    pairs = n * (n - 1) // 2
    for _ in range(args.warmup):
        cpu_search(orc, season, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = cpu_search(orc, season, threads)
    dt = time.perf_counter() - t0
    value = pairs * args.steps / dt
    # the fingerprint stage beside it: ~20 s of oracle work
    made = make_pcm_segments(range(4), 1, SEASON_MINUTES)
    segs = [made[v][k] for v in range(4) for k in (0, 1)]
    t0 = time.perf_counter()
    cpu_fingerprint(orc, segs, threads)
    t_fp = time.perf_counter() - t0
    hours = sum(s.size for s in segs) / synth.SAMPLE_RATE / 3600.0
    sample = ("BOUNDED SAMPLE: every step matches the first %d of the 200 episodes (%d of 19,900 pairs, openings + "
              "endings) and votes; CPU restatement of the reference algorithm in C (oracle/match_ref.c), one worker "
              "thread per pair like the reference's rayon par_iter (comparator.rs:549-564); the Rust reference cannot "
              "be built here (no cargo)" % (n, pairs))
    line = {
        "impl": "reference", "metric": "episode_pairs_per_sec", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "search-only from precomputed hashes, 200 episodes x 24 min, 19,900 pairs, openings + "
                               "endings (BASELINE configs[3])", "episodes": n, "pairs": pairs,
                   "cells": season_cells(season)},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fingerprint": {"metric": "audio_hours_per_sec", "value": hours / t_fp, "unit": "audio-hours/s",
                        "sample": "%.2f audio-hours (8 segments of the configs[1] season) once, oracle/chromaprint_ref.c, "
                                  "%d threads, %.1f s" % (hours, threads, t_fp)},
        "openings_found": int(sum(r[1] for r in res)), "endings_found": int(sum(r[2] for r in res)),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ b200 arm

def run_b200(args):
    # a stuck collective must not hold the GPU box until the driver's limit: dump every thread's stack and leave
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ.get("NB200_BENCH_WATCHDOG_S", "600")), exit=True)
    # NCCL's banner and warnings go to stderr: stdout carries one JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    from needle_b200 import engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    legs = set(args.legs.split(","))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    assert world == max(args.gpus, 1) or world == 1, "launch with torchrun --nproc-per-node = --gpus"

    ctx = engine.Context(local_rank)
    # the library's kernels run on this stream, so that the CUDA events below bracket them
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    # the multi-GPU jobs live behind the C ABI: torch.distributed only carries NCCL's unique id, the
    # barrier and the max-over-ranks of the timings
    from needle_b200 import dist as nd
    comm = nd.comm_from_torch(ctx, dist, device=dev)
    params = engine.match_params(include_endings=True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, steps, warmup):
        """W warm-up steps, then K steps between barriers, timed on the device with CUDA events on
        the stream the library launches on; max over ranks.  -> ms per step"""
        for _ in range(warmup):
            fn()
        barrier()
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1)) / steps

    W = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    line = {}
    parity = {}

    # ============================================================ search (configs[3]) -- the headline
    season = search_season()
    n_pairs = SEARCH_EPISODES * (SEARCH_EPISODES - 1) // 2
    cells = season_cells(season)
    job = engine.MultiJob.search([comm], season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, params)
    results = job.run()
    for _ in range(int(os.environ.get("NB200_BENCH_SPIN_STEPS", "60"))):   # clocks up before anything is timed
        job.run()
    launches0 = ctx.last_kernel_ms()["n_launches"]
    sampler.active = True
    phase_acc = {}
    kernel_acc = {"match": 0.0, "simhash": 0.0, "vote": 0.0}

    search_ms = timed(job.run, args.steps, W)      # nothing but the job call inside the timed region
    sampler.active = False
    launches = (ctx.last_kernel_ms()["n_launches"] - launches0) // (args.steps + W) * args.steps
    n_calls = 5                                    # per-phase CUDA-event times from five more, untimed, steps
    for _ in range(n_calls):
        job.run()
        for k, v in job.phase_ms().items():
            phase_acc[k] = phase_acc.get(k, 0.0) + v
        km = ctx.last_kernel_ms()
        for k in kernel_acc:
            kernel_acc[k] += km[k]
    phase_ms = {k: v / n_calls for k, v in phase_acc.items()}
    kernel_ms = {k: v / n_calls for k, v in kernel_acc.items()}
    match_ms_max = max_over_ranks(kernel_ms["match"])

    # host buffers in, results out, every step: page-locked arrays (the library copies straight from them),
    # and, reported beside it, ordinary pageable ones (staged through the library's own pinned area)
    pin_h = engine.PinnedArray.empty(season.hashes.size, np.uint32)
    pin_t = engine.PinnedArray.empty(season.ts_ns.size, np.uint64)
    pin_h.array[:] = season.hashes
    pin_t.array[:] = season.ts_ns

    def search_e2e_step(h=pin_h.array, t=pin_t.array):
        if world == 1:
            return ctx.search(h, t, season.seg_offset, season.hash_duration_ns, params)
        j = engine.MultiJob.search([comm], h, t, season.seg_offset, season.hash_duration_ns, params)
        r = j.run()
        j.free()
        return r

    e2e_results = search_e2e_step()
    sampler.active = True
    search_e2e_ms = timed(search_e2e_step, args.steps, W)
    sampler.active = False
    search_e2e_pageable_ms = timed(lambda: search_e2e_step(season.hashes, season.ts_ns), args.steps, W)
    if rank == 0:
        assert e2e_results == results, "host-buffer search and resident search disagree"
        # N-rank job against the single-GPU call on the whole season
        with engine.Context(local_rank) as c1:
            single = c1.search(season.hashes, season.ts_ns, season.seg_offset, season.hash_duration_ns, params)
        parity["search_multi_vs_single_gpu"] = {"videos": SEARCH_EPISODES,
                                                "agreeing": sum(1 for a, b in zip(results, single) if a == b)}

    # the exhaustive match kernel (one POPC per cell: the kernel the POPC roofline is about), live, this rank's slice
    from needle_b200._lib import OPT_MATCH_DENSE
    ctx.set_option(OPT_MATCH_DENSE, 1)
    job.run()
    dense_ms = None
    for _ in range(3):
        job.run()
        m = ctx.last_kernel_ms()["match"]
        dense_ms = m if dense_ms is None else min(dense_ms, m)
    ctx.set_option(OPT_MATCH_DENSE, 0)
    dense_ms = max_over_ranks(dense_ms)
    job.free()

    # ============================================================ fingerprint (configs[4])
    fp = None
    if "fingerprint" in legs:
        ep_n = int(round(FP_EPISODE_MIN * 60.0 * synth.SAMPLE_RATE))
        a0, b0, _seek = synth.split_segments(np.zeros(ep_n, np.int16))
        seg_n = [a0.size, b0.size]
        hours_per_episode = (a0.size + b0.size) / synth.SAMPLE_RATE / 3600.0
        n_episodes = int(np.ceil(args.fp_hours / hours_per_episode))
        my_eps = list(range(rank, n_episodes, world))          # episodes sharded round-robin over the ranks
        made = make_pcm_segments(range(FP_POOL), 7, FP_EPISODE_MIN)
        pool = np.concatenate([np.concatenate([made[v][0], np.zeros((-made[v][0].size) % 8, np.int16),
                                               made[v][1], np.zeros((-made[v][1].size) % 8, np.int16)])
                               for v in range(FP_POOL)])
        ep_stride = pool.size // FP_POOL
        end_off = (a0.size + 7) // 8 * 8
        n_tiles = (len(my_eps) + FP_POOL - 1) // FP_POOL
        d_pool = torch.from_numpy(pool).to(dev)
        d_pcm = torch.empty(n_tiles * pool.size + 64, dtype=torch.int16, device=dev)
        d_pcm[:n_tiles * pool.size].view(n_tiles, pool.size).copy_(d_pool.unsqueeze(0).expand(n_tiles, pool.size))
        offs, cnts = [], []
        for k in range(len(my_eps)):
            base = k * ep_stride
            offs += [base, base + end_off]
            cnts += seg_n
        torch.cuda.synchronize(dev)
        ps = engine.PcmSet.view(ctx, d_pcm.data_ptr(), offs, cnts, d_pcm.numel(), keepalive=d_pcm)
        frames_local = sum(synth.num_frames(c) for c in cnts)
        hours_local = sum(cnts) / synth.SAMPLE_RATE / 3600.0
        fp_k = {"fp_fft_chroma": 0.0, "fp_classify": 0.0}

        def fp_step():
            hs = ps.fingerprint(stride=2)
            km = ctx.last_kernel_ms()
            for k in fp_k:
                fp_k[k] += km[k]
            hs.free()

        fp_steps = max(3, min(args.steps, 10))
        sampler.active = True
        fp_ms = timed(fp_step, fp_steps, 3)
        sampler.active = False
        k1_ms = fp_k["fp_fft_chroma"] / (fp_steps + 3)
        k2_ms = fp_k["fp_classify"] / (fp_steps + 3)
        hours_total = sum_over_ranks(hours_local)
        frames_total = sum_over_ranks(frames_local)
        # parity of the tiled job: every copy of a pool episode must hash like the first
        hs = ps.fingerprint(stride=2)
        h, _t, off = hs.download()
        hs.free()
        same = all(np.array_equal(h[off[2 * k]:off[2 * k + 2]], h[off[2 * (k % FP_POOL)]:off[2 * (k % FP_POOL) + 2]])
                   for k in range(FP_POOL, len(my_eps), max(1, len(my_eps) // 97)))
        ps.free()
        del d_pcm
        torch.cuda.empty_cache()
        # e2e sample: pinned host PCM of the pool, H2D pipelined under K1, hashes back to the host
        pin = engine.PinnedArray.empty(pool.size, np.int16)
        pin.array[:] = pool
        host_segs, hcnt = [], []
        for v in range(FP_POOL):
            base = v * ep_stride
            host_segs += [pin.array[base:base + seg_n[0]], pin.array[base + end_off:base + end_off + seg_n[1]]]
        _o, _l, cap = nd.device_layout([s.size for s in host_segs], 2)
        d_h = torch.zeros(cap, dtype=torch.int32, device=dev)
        d_t = torch.zeros(cap, dtype=torch.int64, device=dev)
        h_out = torch.empty(cap, dtype=torch.int32).pin_memory()

        def fp_e2e_step():
            ctx.fingerprint_host_into(host_segs, d_h.data_ptr(), d_t.data_ptr(), cap, stride=2)
            ctx.synchronize()
            h_out.copy_(d_h, non_blocking=True)
            torch.cuda.synchronize(dev)

        fp_e2e_ms = timed(fp_e2e_step, fp_steps, 3)
        e2e_hours = world * sum(s.size for s in host_segs) / synth.SAMPLE_RATE / 3600.0
        pin.free()
        fp = {"ms": fp_ms, "k1_ms": k1_ms, "k2_ms": k2_ms, "hours": hours_total, "frames": frames_total,
              "frames_local": frames_local, "episodes": n_episodes, "e2e_ms": fp_e2e_ms, "e2e_hours": e2e_hours,
              "e2e_h2d": int(sum(s.size for s in host_segs) * 2), "e2e_d2h": int(cap * 4), "steps": fp_steps,
              "tiles_equal": bool(same)}

    # ============================================================ season (configs[1]): analyze + search, host PCM
    se = None
    if "season" in legs:
        n_mono, seeks = segment_metadata(SEASON_EPISODES, SEASON_MINUTES)
        sjob = engine.MultiJob.season([comm], n_mono, seeks, synth.HASH_DURATION_NS, params)
        vr = sjob.video_rank()
        my_videos = [v for v in range(SEASON_EPISODES) if vr[v] == rank]
        made = make_pcm_segments(my_videos, 1, SEASON_MINUTES)
        total = sum(made[v][0].size + made[v][1].size for v in my_videos)
        pinned = engine.PinnedArray.empty(total, np.int16)
        mine, pos = {}, 0
        for v in my_videos:
            for k in (0, 1):
                x = made[v][k]
                pinned.array[pos:pos + x.size] = x
                mine[2 * v + k] = pinned.array[pos:pos + x.size]
                pos += x.size
        del made
        season_results = sjob.run(mine)
        sampler.active = True
        if world == 1:
            flat = [mine[s] for s in range(2 * SEASON_EPISODES)]
            one_call = ctx.analyze_search(flat, 1, seeks, synth.HASH_DURATION_NS, params)
            assert one_call == season_results, "nb200_analyze_search and the 1-rank season job disagree"
            season_ms = timed(lambda: ctx.analyze_search(flat, 1, seeks, synth.HASH_DURATION_NS, params), args.steps, W)
        else:
            season_ms = timed(lambda: sjob.run(mine), args.steps, W)
        sampler.active = False
        season_phase = sjob.phase_ms()
        # resident PCM: the same job without the host copies
        sjob.upload_pcm(mine)
        res_resident = sjob.run()
        season_res_ms = timed(lambda: sjob.run(), args.steps, W)
        if rank == 0:
            assert res_resident == season_results, "host-PCM path and resident path disagree"
        sjob.free()
        se = {"ms": season_ms, "resident_ms": season_res_ms, "h2d": max_over_ranks(total * 2), "h2d_total": sum_over_ranks(total * 2),
              "phase": season_phase, "results": season_results, "mine": mine, "pinned": pinned, "seeks": seeks}

    clocks = sampler.summary()

    # ============================================================ CPU baseline + parity vs the oracle (rank 0)
    cpu = None
    if rank == 0 and args.cpu_baseline:
        from oracle import oracle as orc
        orc.lib()
        threads = os.cpu_count() or 1
        n_sub = 40
        sub = sub_season(season, n_sub)
        t0 = time.perf_counter()
        ref = cpu_search(orc, sub, threads)
        t_cpu = time.perf_counter() - t0
        with engine.Context(local_rank) as c1:
            got = c1.search(sub.hashes, sub.ts_ns, sub.seg_offset, sub.hash_duration_ns, params)
        parity["search_gpu_vs_oracle"] = {"videos": n_sub, "agreeing": sum(1 for a, b in zip(got, ref) if tuple(a) == tuple(b)),
                                          "bar": "bit-exact"}
        sub_pairs = n_sub * (n_sub - 1) // 2
        cpu = {"value": sub_pairs / t_cpu, "unit": "pairs/s", "cores": threads, "kind": "port",
               "sample": "the first %d of the 200 episodes once (%d of 19,900 pairs, openings + endings, vote): CPU "
                         "restatement of the reference algorithm (oracle/match_ref.c, one worker thread per pair), "
                         "%.2f s" % (n_sub, sub_pairs, t_cpu)}
        if se is not None and world == 1:
            # the analyze + search path against the oracle on the first 6 episodes of the configs[1] season
            nv = 6
            segs6 = [se["mine"][s] for s in range(2 * nv)]
            t0 = time.perf_counter()
            raw = cpu_fingerprint(orc, segs6, threads)
            t_fp = time.perf_counter() - t0
            hs_, ts_, off_ = [], [], [0]
            for s in range(2 * nv):
                h_, t_ = orc.subsample_and_stamp(raw[s], 2, seek_to_ns=int(se["seeks"][s]))
                hs_.append(h_)
                ts_.append(t_)
                off_.append(off_[-1] + h_.size)
            s6 = orc.Season(np.concatenate(hs_), np.concatenate(ts_), np.asarray(off_, np.uint64),
                            np.full(nv, synth.HASH_DURATION_NS, np.uint64))
            st, ref6, _ = orc.run_with_frame_hashes(s6, include_endings=True, n_threads=threads)
            with engine.Context(local_rank) as c1:
                got6 = c1.analyze_search(segs6, 1, se["seeks"][:2 * nv], synth.HASH_DURATION_NS, params)
            tol = 2 * 123_000_000    # one hash period (north_star)
            agree = sum(1 for g, w in zip(got6, ref6)
                        if tuple(g[:3]) == tuple(int(x) for x in w[:3]) and
                        all(abs(int(a) - int(b)) <= tol for a, b in zip(g[3:], w[3:])))
            parity["season_gpu_vs_oracle"] = {"videos": nv, "agreeing": agree, "bar": "intervals within one hash period"}
            allsegs = [se["mine"][s] for s in range(2 * SEASON_EPISODES)]
            t0 = time.perf_counter()
            for _ in range(3):
                cpu_fingerprint(orc, allsegs, threads)
            t_fp = (time.perf_counter() - t0) / 3
            hrs = sum(x.size for x in allsegs) / synth.SAMPLE_RATE / 3600.0
            cpu["fingerprint"] = {"value": hrs / t_fp, "unit": "audio-hours/s",
                                  "sample": "%.1f audio-hours (the 56 segments of the configs[1] season) three times, "
                                            "oracle/chromaprint_ref.c, one worker thread per segment like the reference's "
                                            "rayon loop over videos, %d threads, %.2f s per pass" % (hrs, threads, t_fp)}
    if rank == 0 and se is not None and world > 1:
        # N-rank season job against the single-GPU call: rank 0 regenerates the whole season
        made = make_pcm_segments(range(SEASON_EPISODES), 1, SEASON_MINUTES)
        flat = [made[v][k] for v in range(SEASON_EPISODES) for k in (0, 1)]
        with engine.Context(local_rank) as c1:
            single = c1.analyze_search(flat, 1, se["seeks"], synth.HASH_DURATION_NS, params)
        parity["season_multi_vs_single_gpu"] = {"videos": SEASON_EPISODES,
                                                "agreeing": sum(1 for a, b in zip(se["results"], single) if a == b)}

    if rank != 0:
        comm.destroy()
        if dist is not None:
            dist.destroy_process_group()
        return

    # ============================================================ the line
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    sm_mhz = clocks["sm_mhz"] or SM_MHZ_MAX
    lane_ops, lane_src = FP32_LANE_OPS_PER_CLK_PER_SM, "profiles/r02_k1_fp_mix_peak.jsonl"
    try:   # the measurement for the current kernel's mix, when it has been taken on this pool
        for ln in open(os.path.join(ROOT, "profiles", "r02_k1_fp_mix_peak_v4.jsonl")):
            rec = json.loads(ln)
            if rec.get("test", "").startswith("k1 fp mix, revision-4"):
                lane_ops, lane_src = float(rec["fp32_lane_ops_per_clk_per_sm"]), "profiles/r02_k1_fp_mix_peak_v4.jsonl"
    except Exception:
        pass
    fp32_peak = SM_COUNT * lane_ops * 2 * SM_MHZ_MAX * 1e6 / 1e12
    popc_peak = SM_COUNT * POPC_PER_CLK_PER_SM * SM_MHZ_MAX * 1e6 / 1e12
    traffic_per_frame, traffic_src = None, None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_k1_traffic.json")))
        traffic_per_frame = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["frames"]
        traffic_src = "profiles/r02_ncu_k1_traffic.json (ncu --set full of this kernel on %d frames, scaled per frame)" % t["frames"]
    except Exception:
        pass

    line = {
        "metric": "episode_pairs_per_sec", "value": n_pairs / (search_ms * 1e-3), "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": search_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 (match), f32+f64 (fingerprint)",
        "data": "synthetic",
        "config": {
            "workload": "search-only from precomputed hashes: 200 episodes x 24 min, 19,900 pairs, openings + endings "
                        "(BASELINE configs[3]); the pair list is sharded over the GPUs by table cells, runs pushed to "
                        "rank 0 over NVLink peer memory, device vote on rank 0 (nb200_mjob_search_*)",
            "episodes": SEARCH_EPISODES, "pairs": n_pairs, "cells": cells,
            "l2": "the season (10.4 MB) is L2-resident by design: the match is POPC-bound, not memory-bound; "
                  "the fingerprint leg reads %.1f GB per GPU per step (larger than L2)" %
                  ((fp["frames_local"] * 2730 / 1e9) if fp else 0.0),
        },
        "e2e": {"value": n_pairs / (search_e2e_ms * 1e-3), "unit": "pairs/s",
                "h2d_bytes_per_step": int(season.hashes.nbytes + season.ts_ns.nbytes) * world,
                "d2h_bytes_per_step": 48 * SEARCH_EPISODES + 64, "ms_per_step": search_e2e_ms,
                "call": "nb200_search (C ABI: page-locked host hash + timestamp arrays in, per-video results out)" if world == 1 else
                        "nb200_mjob_search_create + nb200_mjob_run + nb200_mjob_free per step on every rank "
                        "(each rank uploads the season from page-locked host arrays)",
                "pageable_input": {"value": n_pairs / (search_e2e_pageable_ms * 1e-3), "ms_per_step": search_e2e_pageable_ms,
                                   "note": "same call on ordinary (pageable) numpy arrays: staged through the library's pinned area"}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "kernel_ms_per_step_rank0": kernel_ms,
        "collective_ms": {"match_slowest_rank": match_ms_max, "run_push_wait_rank0": phase_ms.get("run_push_wait", 0.0),
                          "vote_rank0": phase_ms.get("vote", 0.0),
                          "note": "CUDA events on rank 0's stream (nb200_mjob_phase_ms): the run blocks travel by a push "
                                  "kernel over NVLink peer memory, not by a collective; run_push_wait includes waiting for "
                                  "the slowest rank's match"},
        "parity": parity,
        "openings_found": int(sum(r[1] for r in results)), "endings_found": int(sum(r[2] for r in results)),
    }
    line["roofline_popc"] = {
        "kernel": "match_fast_kernel<dense> (NB200_OPT_MATCH_DENSE=1): one POPC per cell", "bound": "int_popc",
        "unit": "Tcell/s", "achieved": cells / world / (dense_ms * 1e-3) / 1e12,
        "kernel_ms": dense_ms, "cells_per_launch": cells // world,
        "peak": popc_peak, "peak_source": "MEASURED POPC issue rate %.1f/clk/SM (tools/pipe_peak.cu, "
                                          "profiles/r01_pipe_peak_warm.jsonl; nominal 16) x 148 SM x 1965 MHz" % POPC_PER_CLK_PER_SM,
        "default_kernel": "match_fast_kernel<adaptive>: a first look at 3 rows of each 32-row word (2 POPCs through a carry-save "
                          "adder), then 4 rows per stage; leaves a word when no diagonal survives -- ~0.08 POPCs per cell on "
                          "unrelated hashes, identical runs (tests/test_match_gpu.py)",
        "default_kernel_xu_pipe_pct": ncu_metric("profiles/r02_ncu_k3_final.txt", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "default_kernel_ms": match_ms_max,
        "default_kernel_algorithmic_Tcell_per_s": cells / world / (match_ms_max * 1e-3) / 1e12,
    }
    line["roofline_popc"]["frac"] = line["roofline_popc"]["achieved"] / popc_peak
    if fp is not None:
        k1_s = fp["k1_ms"] * 1e-3
        achieved = fp["frames_local"] * FLOP_PER_FRAME / k1_s / 1e12
        line["roofline"] = {
            "kernel": "fp_fft_chroma_tm_kernel (K1) of the fingerprint leg: the kernel with the largest share of this run's GPU time",
            "bound": "fp32", "unit": "TFLOP/s", "achieved": achieved, "peak": fp32_peak, "frac": achieved / fp32_peak,
            "peak_source": "MEASURED: K1's own FP32 instruction mix with no memory or integer work reaches %.1f results per "
                           "clock per SM (tools/k1_mix_peak.cu, %s; nominal 128) x 2 flop "
                           "x 148 SM x 1965 MHz" % (lane_ops, lane_src),
            "frac_of_nominal_74p4": achieved / (SM_COUNT * 128 * 2 * SM_MHZ_MAX * 1e6 / 1e12),
            "frames_per_launch": fp["frames_local"], "flop_per_frame": FLOP_PER_FRAME, "kernel_ms": fp["k1_ms"],
            "traffic": (traffic_per_frame * fp["frames_local"]) if traffic_per_frame else None, "traffic_source": traffic_src,
            "hbm": {"achieved": fp["frames_local"] * BYTES_PER_FRAME / k1_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": fp["frames_local"] * BYTES_PER_FRAME / k1_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                    "note": "secondary: 49 flop/B, the kernel is FP32-bound (SURVEY.md 8d)"},
        }
        line["fingerprint"] = {
            "metric": "audio_hours_per_sec", "value": fp["hours"] / (fp["ms"] * 1e-3), "unit": "audio-hours/s",
            "ms_per_step": fp["ms"], "steps": fp["steps"], "scaling": "strong",
            "config": {"workload": "fingerprint-only, %.0f audio-hours of 11025 Hz mono PCM resident in HBM (BASELINE "
                                   "configs[4]): %d episodes x 24 min, opening 50 %% + ending 25 %% of each, episodes "
                                   "sharded over the GPUs; %d distinct synthetic episodes tiled on the device" %
                                   (fp["hours"], fp["episodes"], FP_POOL),
                       "audio_hours": fp["hours"], "frames": fp["frames"], "episodes": fp["episodes"]},
            "kernel_ms_per_step_rank0": {"fp_fft_chroma": fp["k1_ms"], "fp_classify": fp["k2_ms"]},
            "tiled_copies_hash_identically": fp["tiles_equal"],
            "e2e": {"value": fp["e2e_hours"] / (fp["e2e_ms"] * 1e-3), "unit": "audio-hours/s", "ms_per_step": fp["e2e_ms"],
                    "h2d_bytes_per_step": fp["e2e_h2d"] * world, "d2h_bytes_per_step": fp["e2e_d2h"] * world,
                    "sample": "BOUNDED SAMPLE: %.1f audio-hours per GPU per step from pinned host memory "
                              "(nb200_fingerprint_host_into: H2D pipelined under K1), hashes copied back" %
                              (fp["e2e_hours"] / world)},
        }
    if se is not None:
        sp = SEASON_EPISODES * (SEASON_EPISODES - 1) // 2
        h2d_peak, h2d_src = None, None
        try:
            hb = json.load(open(os.path.join(ROOT, "profiles", "r02_h2d_bandwidth_n.json")))
            h2d_peak = float(hb[str(world)]["aggregate_GBs"])
            h2d_src = "profiles/r02_h2d_bandwidth_n.json (tools/h2d_bw.py: %d concurrent pinned copies)" % world
        except Exception:
            pass
        ach = se["h2d_total"] / (se["ms"] * 1e-3) / 1e9
        line["season_e2e"] = {
            "metric": "episode_pairs_per_sec", "value": sp / (se["ms"] * 1e-3), "unit": "pairs/s", "ms_per_step": se["ms"],
            "h2d_bytes_per_step": int(se["h2d_total"]), "d2h_bytes_per_step": 48 * SEASON_EPISODES + 64,
            "resident_value": sp / (se["resident_ms"] * 1e-3), "resident_ms_per_step": se["resident_ms"],
            "config": {"workload": "28 episodes x 20 min, analyze + search with endings (BASELINE configs[1]) from pinned "
                                   "host PCM; episodes sharded for fingerprinting, ONE ncclAllGather of the hashes, pairs "
                                   "sharded for matching, runs pushed to rank 0, device vote", "pairs": sp,
                       "audio_hours_fingerprinted": se["h2d_total"] / 2 / synth.SAMPLE_RATE / 3600.0},
            "call": "nb200_analyze_search (C ABI, pinned host PCM)" if world == 1 else "nb200_mjob_run (season job, host PCM)",
            "phase_ms_rank0": se["phase"],
            "roofline": {"bound": "pcie", "unit": "GB/s", "achieved": ach, "peak": h2d_peak,
                         "frac": (ach / h2d_peak) if h2d_peak else None, "peak_source": h2d_src},
            "openings_found": int(sum(r[1] for r in se["results"])), "endings_found": int(sum(r[2] for r in se["results"])),
        }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    comm.destroy()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
