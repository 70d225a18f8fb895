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

# The same with pitched 2-D copies of the shape pm_upload_slab / pm_download_slab use at 8192^2
# (dense host rows of 8193 doubles <-> device rows of pitch 8224 doubles).
from cuda.bindings import runtime as rt  # noqa: E402

rows, cols, pitch = 8194, 8193, 8224
hp_in, hp_out = h_in.data_ptr(), h_out.data_ptr()
dp_in, dp_out = d_in.data_ptr(), d_out.data_ptr()
K = rt.cudaMemcpyKind


def run2d(up, down, reps=4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            rt.cudaMemcpy2DAsync(dp_in, pitch * 8, hp_in, cols * 8, cols * 8, rows, K.cudaMemcpyHostToDevice, s1.cuda_stream)
        if down:
            rt.cudaMemcpy2DAsync(hp_out, cols * 8, dp_out, pitch * 8, cols * 8, rows, K.cudaMemcpyDeviceToHost, s2.cuda_stream)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * rows * cols * 8 / dt / 1e9


run2d(True, True, 1)
print(f"2-D pitched: H2D alone {run2d(True, False):.1f} GB/s; D2H alone {run2d(False, True):.1f} GB/s; both: {run2d(True, True):.1f} GB/s each direction")
