"""H2D alone, D2H alone and both at once (two streams, pinned buffers): does this box overlap the two directions?"""
import time
import torch

n = 134283264 // 16
h_in = torch.empty(n, dtype=torch.complex128).pin_memory()
h_out = torch.empty(n, dtype=torch.complex128).pin_memory()
d_in = torch.empty(n, dtype=torch.complex128, device="cuda")
d_out = torch.empty(n, dtype=torch.complex128, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=10):
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
    return (time.perf_counter() - t0) / reps * 1e3


for _ in range(2):
    run(True, True, 2)
b = h_in.numel() * 16 / 1e9
u, d, both = run(True, False), run(False, True), run(True, True)
print("134 MB: H2D %.2f ms (%.1f GB/s) | D2H %.2f ms (%.1f GB/s) | both at once %.2f ms (%.1f GB/s aggregate)" % (u, b / u * 1e3, d, b / d * 1e3, both, 2 * b / both * 1e3))
