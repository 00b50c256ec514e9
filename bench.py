#!/usr/bin/env python
"""bench.py -- the fingerprint-and-match path on the BASELINE.json workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): a 28-episode synthetic season, 20 min per
episode, `analyze --include-endings` + `search --include-endings` = fingerprint
the first 50 % and last 25 % of every episode, match all 378 pairs (openings and
endings), vote.  One step = one pass of that path over one season per GPU.
With N GPUs the job is a library of N such seasons (weak scaling): episodes are
sharded over ranks for fingerprinting, the hash arrays are all-gathered once
(NCCL), the N*378 within-season pairs are sharded over ranks for matching, every
rank's runs land in one device block, the blocks are all-gathered once and rank
0's GPU votes.

Prints ONE JSON line (rank 0).  `value` = episode-pairs/s with the PCM already
resident in HBM; `e2e` = the same through the host-buffer C-ABI call (pinned
host PCM -> H2D -> kernels incl. the vote -> D2H of the result table).  `--impl reference` times
the CPU restatement of the reference path (oracle/, all host threads): the
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

EPISODES = 28
MINUTES = 20.0
FLOP_PER_FRAME = 134.6e3        # 4096-pt real FFT + window + power + fold + classify
BYTES_PER_FRAME = 1365 * 2 + 4  # mono i16 in (one hop) + one u32 hash out
POPC_PER_CLK_PER_SM = 16.0      # XU pipe; tools/pipe_peak.cu measures 15.0 at 1965 MHz (profiles/r01_pipe_peak_warm.jsonl)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--episodes", type=int, default=EPISODES)
    ap.add_argument("--minutes", type=float, default=MINUTES)
    ap.add_argument("--cpu-baseline", type=int, default=1, help="time the CPU oracle beside the GPU run (N=1)")
    return ap.parse_args()


# ------------------------------------------------------------------ workload

def make_segments(video_ids, episodes_per_season, minutes):
    """PCM of the given global video ids (video v = episode v % E of season v // E).
    Returns {video: (opening_pcm, ending_pcm, ending_seek_ns)}."""
    themes = {}

    def one(v):
        season = 1 + v // episodes_per_season
        if season not in themes:
            themes[season] = synth.season_themes(season)
        ep = synth.make_pcm_episode(season, v % episodes_per_season, minutes, *themes[season])
        return v, synth.split_segments(ep.pcm)

    for v in video_ids:    # build the themes serially (cheap), episodes in parallel
        s = 1 + v // episodes_per_season
        if s not in themes:
            themes[s] = synth.season_themes(s)
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:
        return dict(ex.map(one, video_ids))


def segment_metadata(n_videos, minutes):
    """Sample counts and seeks of every segment, computable without the audio."""
    n = int(round(minutes * 60.0 * synth.SAMPLE_RATE))
    a, b, seek = synth.split_segments(np.zeros(n, np.int16))
    n_mono = [a.size, b.size] * n_videos
    seeks = [0, seek] * n_videos
    return n_mono, seeks


def within_season_pairs(n_seasons, per_season):
    out = []
    for s in range(n_seasons):
        base = s * per_season
        out += [(base + i, base + j) for i in range(per_season) for j in range(i + 1, per_season)]
    return np.array(out, dtype=np.uint32).reshape(-1, 2)


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

def cpu_reference_step(orc, segs_by_video, seeks, n_videos, per_season, threads):
    """One pass of the path on the CPU restatement: Analyzer::run over the
    videos (one worker per video, analyzer.rs:437-445), then
    Comparator::run_with_frame_hashes per season (one worker per pair)."""
    flat = []
    for v in range(n_videos):
        flat += [segs_by_video[v][0], segs_by_video[v][1]]
    raw = orc.fingerprint_many(flat, channels=1, n_threads=threads)
    results = []
    for s in range(n_videos // per_season):
        op, en = [], []
        for v in range(s * per_season, (s + 1) * per_season):
            op.append(orc.subsample_and_stamp(raw[2 * v], 2))
            en.append(orc.subsample_and_stamp(raw[2 * v + 1], 2, seek_to_ns=int(seeks[2 * v + 1])))
        hs, ts, off = [], [], [0]
        for (oh, ot), (eh, et) in zip(op, en):
            hs += [oh, eh]
            ts += [ot, et]
            off += [off[-1] + oh.size, off[-1] + oh.size + eh.size]
        season = orc.Season(np.concatenate(hs), np.concatenate(ts), np.asarray(off, np.uint64),
                            np.full(per_season, synth.HASH_DURATION_NS, np.uint64))
        st, res, _ = orc.run_with_frame_hashes(season, include_endings=True, n_threads=threads)
        assert st == 0
        results += res
    return results


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.lib()
    threads = os.cpu_count() or 1
    n_videos = args.episodes
    segs = make_segments(range(n_videos), args.episodes, args.minutes)
    _, seeks = segment_metadata(n_videos, args.minutes)
    # Bounded run: one probe step on the full season; if K + W of those would not end
    # within a few minutes, every step works on the first n' episodes instead (stated in `sample`).
    t0 = time.perf_counter()
    cpu_reference_step(orc, segs, seeks, n_videos, n_videos, threads)
    probe = time.perf_counter() - t0
    budget_s = 150.0
    full = n_videos
    while n_videos > 4 and probe * (n_videos / full) ** 1.5 * (args.steps + args.warmup) > budget_s:
        n_videos -= 2
    sampled = n_videos != full
    pairs = n_videos * (n_videos - 1) // 2
    for _ in range(args.warmup):
        cpu_reference_step(orc, segs, seeks, n_videos, n_videos, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = cpu_reference_step(orc, segs, seeks, n_videos, n_videos, threads)
    dt = time.perf_counter() - t0
    value = pairs * args.steps / dt
    hours = sum(segs[v][0].size + segs[v][1].size for v in range(n_videos)) / synth.SAMPLE_RATE / 3600.0
    line = {
        "impl": "reference", "metric": "episode_pairs_per_sec", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+u32", "data": "synthetic",
        "config": {"workload": "28x20min season, analyze+search with endings (BASELINE configs[1])",
                   "episodes": n_videos, "minutes": args.minutes, "pairs": pairs, "audio_hours_fingerprinted": hours},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": ("BOUNDED SAMPLE (first episodes of the season, a full step takes %.1f s here): " % probe if sampled
                                    else "full workload per step: ") +
                                   "%d episodes fingerprinted (opening 50%% + ending 25%%), "
                                   "%d pairs matched, vote; CPU restatement of the reference algorithm in C "
                                   "(oracle/), one worker thread per video / per pair like the reference's rayon "
                                   "par_iter; the Rust reference cannot be built here (no cargo)" % (n_videos, pairs)},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "audio_hours_per_sec": hours * args.steps / dt,
        "openings_found": int(sum(r[1] for r in res)), "endings_found": int(sum(r[2] for r in res)),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ b200 arm

def run_b200(args):
    # a stuck collective must not hold the GPU box until the driver's limit: dump every thread's
    # stack and leave
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ.get("NB200_BENCH_WATCHDOG_S", "420")), exit=True)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"    # NCCL's version banner goes to stdout; this run prints one JSON line
    import torch
    from needle_b200 import dist as nd
    from needle_b200 import engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == max(args.gpus, 1) or world == 1, "launch with torchrun --nproc-per-node = --gpus"

    # N > 1: every rank streams 556 MB of pinned PCM per step; keep each rank's host memory on
    # its GPU's NUMA node (N = 1 keeps all cores: the CPU baseline runs in this process)
    numa_bound = nd.bind_host_to_gpu_numa(local_rank) if world > 1 else False
    per_season = args.episodes
    n_videos = per_season * world
    backend = nd.GpuBackend(local_rank)
    ctx = backend.ctx
    dev = backend.device
    params = engine.match_params(include_endings=True)
    n_mono, seeks = segment_metadata(n_videos, args.minutes)
    pairs = within_season_pairs(world, per_season)
    hd = np.full(n_videos, synth.HASH_DURATION_NS, np.uint64)
    job = nd.SeasonJob(backend, dist, n_mono, seeks, hd, params, pairs=pairs)

    # this rank's PCM, in pinned host memory
    seg_ids = job.local_segment_ids()
    my_videos = sorted({s // 2 for s in seg_ids})
    made = make_segments(my_videos, per_season, args.minutes)
    total = sum(made[s // 2][s % 2].size for s in seg_ids)
    pinned = engine.PinnedArray.empty(total, np.int16)
    local_segments, pos = [], 0
    for s in seg_ids:
        x = made[s // 2][s % 2]
        pinned.array[pos:pos + x.size] = x
        local_segments.append(pinned.array[pos:pos + x.size])
        pos += x.size
    del made
    h2d_bytes = total * 2
    frames_local = sum(synth.num_frames(x.size) for x in local_segments)
    my_slice = job.slices[rank]
    my_pairs = pairs[my_slice[0]:my_slice[1]]
    sl = job.plan.seg_len.astype(np.int64)
    cells_local = int(sum(sl[2 * a] * sl[2 * b] + sl[2 * a + 1] * sl[2 * b + 1] for a, b in my_pairs))

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

    pcmset = job.upload(local_segments)
    results = job.run_resident(pcmset)          # also the parity sanity numbers printed below
    # bytes that come back per step: this rank's runs + the timestamp mirror the vote needs
    if world == 1:
        _hs = pcmset.fingerprint(stride=2, seek_to_ns=job.local_seek)
        _rs = _hs.match(params)
        n_runs_local = _rs.count()[0]
        _rs.free()
        _hs.free()
    else:
        _season = backend.season_from_gathered(job._buffers(), job.plan, world)   # filled by the run above
        n_runs_local = backend.match(_season, params, my_pairs).shape[0]
        _season.free()
    if world == 1:
        d2h_bytes = 48 * n_videos + 16 + 16    # the result table, the vote's flags, the match counters
    else:
        d2h_bytes = 64 * n_runs_local + 16     # this rank's runs (they carry their timestamps) + counters

    # bring the clocks up before anything is timed (idle parts sit at 120 MHz).  A FIXED number of
    # steps: every step holds collectives, so all ranks must run the same count (a per-rank
    # time limit would let one rank do one step more than another and misalign them)
    for _ in range(int(os.environ.get("NB200_BENCH_SPIN_STEPS", "500"))):
        job.run_resident(pcmset)

    sampler = ClockSampler(local_rank)
    kernel_ms = {"fp_fft_chroma": 0.0, "fp_classify": 0.0, "match": 0.0, "simhash": 0.0, "vote": 0.0}

    # ---- value: inputs resident in HBM
    for _ in range(max(args.warmup, 3)):
        job.run_resident(pcmset)
    barrier()
    launches0 = ctx.last_kernel_ms()["n_launches"]
    ctx.host_profile(reset=True)
    job.phase_s = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active = True
    ev0.record()
    for _ in range(args.steps):
        job.run_resident(pcmset)
        ms = ctx.last_kernel_ms()
        for k in kernel_ms:
            kernel_ms[k] += ms[k]
    ev1.record()
    barrier()
    sampler.active = False
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = ctx.last_kernel_ms()["n_launches"] - launches0
    host_ms = {k: v / args.steps for k, v in ctx.host_profile(reset=True).items()}
    job_ms = {k: v * 1e3 / args.steps for k, v in job.phase_s.items()}
    for k in kernel_ms:
        kernel_ms[k] /= args.steps

    # ---- the exhaustive match kernel (every cell: the POPC-roofline reference) timed live on
    #      this rank's own 28 episodes, outside the timed regions
    from needle_b200._lib import OPT_MATCH_DENSE
    _hs = pcmset.fingerprint(stride=2, seek_to_ns=job.local_seek)
    dense_ms, dense_cells = None, None
    for opt in (1, 0):
        ctx.set_option(OPT_MATCH_DENSE, opt)
        best = None
        for _ in range(5):
            _rs = _hs.match(params)
            dense_cells = _rs.count()[1]
            _rs.free()
            m = ctx.last_kernel_ms()["match"]
            best = m if best is None else min(best, m)
        if opt:
            dense_ms = best
        else:
            adaptive_ms_local = best
    _hs.free()

    # ---- e2e: host buffers in, results out, through the public call
    backend.release(pcmset)
    one_call = world == 1   # N = 1: the single C-ABI call nb200_analyze_search

    def e2e_step():
        if one_call:
            return ctx.analyze_search(local_segments, 1, job.local_seek, synth.HASH_DURATION_NS, params)
        return job.run_host(local_segments)

    for _ in range(max(args.warmup, 3)):
        e2e_results = e2e_step()
    barrier()
    sampler.active = True
    ev0.record()
    for _ in range(args.steps):
        e2e_results = e2e_step()
    ev1.record()
    barrier()
    sampler.active = False
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.summary()
    if rank == 0:
        assert e2e_results == results, "host-buffer path and resident path disagree"

    # ---- CPU baseline beside it (rank 0, N = 1): the oracle on the same workload
    cpu = None
    if rank == 0 and world == 1 and args.cpu_baseline:
        from oracle import oracle as orc
        orc.lib()
        threads = os.cpu_count() or 1
        segs_by_video = {v: (local_segments[2 * k], local_segments[2 * k + 1]) for k, v in enumerate(my_videos)}
        t0 = time.perf_counter()
        ref = cpu_reference_step(orc, segs_by_video, seeks, n_videos, per_season, threads)
        t_cpu = time.perf_counter() - t0
        tol = 2 * 123_000_000    # one hash period (north_star)
        agree = sum(1 for g, w in zip(results, ref)
                    if g[:3] == w[:3] and all(abs(int(a) - int(b)) <= tol for a, b in zip(g[3:], w[3:])))
        cpu = {"value": len(pairs) / t_cpu, "unit": "pairs/s", "cores": threads, "kind": "port",
               "sample": "the full workload once (%d episodes, %d pairs): CPU restatement of the reference "
                         "algorithm (oracle/, C, one worker thread per video / pair), %.2f s" %
                         (n_videos, len(pairs), t_cpu),
               "videos_agreeing_within_one_hash_period": agree, "videos": n_videos}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    sm_mhz = clocks["sm_mhz"] or 1965.0
    k1_s = kernel_ms["fp_fft_chroma"] * 1e-3
    k3_s = kernel_ms["match"] * 1e-3
    fp32_peak_nominal = 148 * 128 * 2 * 1.965e9 / 1e12
    popc_peak_nominal = 148 * POPC_PER_CLK_PER_SM * 1.965e9 / 1e12
    n_pairs_total = len(pairs)
    hours_total = world * sum(x.size for x in local_segments) / synth.SAMPLE_RATE / 3600.0
    step_ms = dev_ms / args.steps
    dominant = "fp_fft_chroma" if kernel_ms["fp_fft_chroma"] >= kernel_ms["match"] else "match"
    # DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture
    # (dram__bytes_read.sum + dram__bytes_write.sum); only valid for the shapes it was taken on
    traffic, traffic_src = None, None
    if args.episodes == EPISODES and args.minutes == MINUTES:
        if dominant == "fp_fft_chroma":
            traffic, traffic_src = 555.75e6 + 13.21e6, "profiles/r01_ncu_k1_h32_v4.txt"
        else:
            traffic, traffic_src = 0.56e6, "profiles/r01_ncu_k3_match_fast_v2.txt"
    line = {
        "metric": "episode_pairs_per_sec", "value": n_pairs_total / (step_ms * 1e-3), "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64+u32",
        "data": "synthetic",
        "config": {
            "workload": "28x20min season per GPU, analyze+search with endings (BASELINE configs[1]); "
                        "N GPUs = N seasons: episodes sharded for fingerprinting, one all-gather of hashes, "
                        "within-season pairs sharded for matching, one all-gather of run blocks, device vote on rank 0",
            "episodes": n_videos, "minutes": args.minutes, "pairs": n_pairs_total,
            "audio_hours_fingerprinted": hours_total,
            "l2": "inputs larger than L2 (%.0f MB of PCM per GPU per step)" % (h2d_bytes / 1e6),
        },
        "e2e": {"value": n_pairs_total / (e2e_ms / args.steps * 1e-3), "unit": "pairs/s",
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(d2h_bytes),
                "ms_per_step": e2e_ms / args.steps, "host_numa_bound": bool(numa_bound),
                "call": "nb200_analyze_search (C ABI, pinned host PCM)" if one_call else
                        "SeasonJob.run_host (pinned host PCM per rank)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "audio_hours_per_sec": hours_total / (step_ms * 1e-3),
        "kernel_ms_per_step": kernel_ms,
        "host_phase_ms_per_step": host_ms,
        "job_phase_ms_per_step_rank0": job_ms,
        # the schema's roofline object, for the kernel with the largest share of the step
        "roofline": {
            "kernel": dominant, "bound": "hbm",
            "achieved": (frames_local * BYTES_PER_FRAME / k1_s / 1e9) if dominant == "fp_fft_chroma" else
                        (4 * sl.sum() / k3_s / 1e9),
            "peak": hbm_peak, "unit": "GB/s", "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": hbm_src,
            "note": "neither kernel is HBM-bound (SURVEY.md 8d: 49 flop/B and 1 POPC per 4e-4 B); "
                    "the binding rooflines are roofline_fp32 (K1) and roofline_popc (K3) below",
        },
        "roofline_fp32": {
            "kernel": "fp_fft_chroma", "bound": "fp32", "unit": "TFLOP/s",
            "achieved": frames_local * FLOP_PER_FRAME / k1_s / 1e12,
            "peak": fp32_peak_nominal, "peak_source": "nominal 148 SM x 128 lanes x 2 x 1965 MHz",
            "peak_at_measured_clock": 148 * 128 * 2 * sm_mhz * 1e6 / 1e12,
            "frames_per_launch": frames_local, "flop_per_frame": FLOP_PER_FRAME,
        },
        "roofline_popc": {
            "kernel": "match_fast_kernel<dense> (NB200_OPT_MATCH_DENSE=1): one POPC per cell", "bound": "int_popc",
            "unit": "Tcell/s",
            "achieved": dense_cells / (dense_ms * 1e-3) / 1e12,
            "kernel_ms": dense_ms,
            "default_kernel": "match_fast_kernel<adaptive>: tests 4 rows of each 32-row word per stage and leaves "
                              "when no diagonal survives, so it executes fewer POPCs than there are cells; "
                              "identical runs (tests/test_match_gpu.py)",
            "default_kernel_ms": kernel_ms["match"],
            "default_kernel_algorithmic_Tcell_per_s": cells_local / k3_s / 1e12,
            "peak": popc_peak_nominal,
            "peak_source": "POPC issue rate %.0f/clk/SM x 148 SM x 1965 MHz (pure POPC loop measures 4.33 T/s, "
                           "profiles/r01_pipe_peak_warm.jsonl)" % POPC_PER_CLK_PER_SM,
            "peak_at_measured_clock": 148 * POPC_PER_CLK_PER_SM * sm_mhz * 1e6 / 1e12,
            "cells_per_launch": dense_cells,
        },
        "openings_found": int(sum(r[1] for r in results)), "endings_found": int(sum(r[2] for r in results)),
    }
    line["roofline"]["frac"] = line["roofline"]["achieved"] / hbm_peak
    for k in ("roofline_fp32", "roofline_popc"):
        line[k]["frac"] = line[k]["achieved"] / line[k]["peak"]
        line[k]["frac_at_measured_clock"] = line[k]["achieved"] / line[k]["peak_at_measured_clock"]
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
