"""PCIe copy rates on this box: H2D alone, D2H alone, both at once (pinned host memory, two streams)."""
import time
import torch

n = 1 << 27  # doubles: 1 GiB
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * n * 8 / dt / 1e9


run(True, True, 1)
print(f"H2D alone {run(True, False):.1f} GB/s; D2H alone {run(False, True):.1f} GB/s; both: {run(True, True):.1f} GB/s each direction")
