"""Host<->device copy bandwidth of this box (pinned and pageable), for reading the e2e numbers."""
import json, time, torch
n = 512 << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
pin = torch.empty(n, dtype=torch.uint8).pin_memory()
page = torch.empty(n, dtype=torch.uint8)
out = {}
for name, src in (("pinned", pin), ("pageable", page)):
    for _ in range(2):
        dev.copy_(src, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dev.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    out["h2d_%s_GBs" % name] = 5 * n / (time.perf_counter() - t0) / 1e9
t0 = time.perf_counter()
for _ in range(5):
    pin.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
out["d2h_pinned_GBs"] = 5 * n / (time.perf_counter() - t0) / 1e9
print(json.dumps(out))
