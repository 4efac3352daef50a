"""Pinned-memory PCIe bandwidth on the box: H2D alone, D2H alone, both directions at once (the e2e pipeline's ceiling)."""
import torch, time, json
N = 1 << 30
h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
d_in = torch.empty(N, dtype=torch.uint8, device="cuda")
d_out = torch.empty(N, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * N / dt / 1e9
run(True, True, 1)
r = {"h2d_GBs": run(True, False), "d2h_GBs": run(False, True), "bidir_each_GBs": run(True, True)}
r["bidir_total_GBs"] = 2 * r["bidir_each_GBs"]
print(json.dumps(r))
