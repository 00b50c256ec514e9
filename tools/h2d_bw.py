"""Host -> device copy ceiling of this box with N GPUs copying at once (one process per GPU, pinned
host memory, 512 MB per copy): the denominator of the end-to-end numbers, which stream PCM from the host.

    python tools/h2d_bw.py [--gpus 1,2,4,8] > profiles/r02_h2d_bandwidth_n.json

For every N in the list the N processes start their copies together (a barrier), each times its own
copies with CUDA events; `aggregate_GBs` = sum over the processes of bytes / the slowest process's time.
Also reports the device -> host direction for N = 1."""
import argparse
import json
import sys
import time


def worker(rank, n, barrier, out_q, d2h):
    import torch
    torch.cuda.set_device(rank)
    nbytes = 512 << 20
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    pin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    pin.fill_(rank + 1)
    src, dst = (dev, pin) if d2h else (pin, dev)
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    reps = 8
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier.wait()
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    out_q.put((rank, reps * nbytes, e0.elapsed_time(e1) * 1e-3))


def measure(n, d2h=False):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(n), ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, n, barrier, q, d2h)) for r in range(n)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(n)]
    for p in procs:
        p.join()
    res.sort()
    slowest = max(t for _, _, t in res)
    return {"aggregate_GBs": sum(b for _, b, _ in res) / slowest / 1e9,
            "per_gpu_GBs": [b / t / 1e9 for _, b, t in res]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    a = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    out = {"direction": "host (pinned) -> device", "bytes_per_copy": 512 << 20, "gpus_on_box": have}
    for n in [int(x) for x in a.gpus.split(",")]:
        if n <= have:
            out[str(n)] = measure(n)
    out["d2h_1"] = measure(1, d2h=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
