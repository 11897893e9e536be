"""Does a kernel on stream B start while a 134 MB D2H (or H2D) copy runs on stream A?"""
import torch
n = 134283264 // 16
h = torch.empty(n, dtype=torch.complex128).pin_memory()
d = torch.empty(n, dtype=torch.complex128, device="cuda")
x = torch.ones(1 << 20, device="cuda")
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
E = lambda: torch.cuda.Event(enable_timing=True)
for kind in ("D2H", "H2D", "D2H", "H2D"):
    torch.cuda.synchronize()
    ref, c0, c1, k0, k1 = E(), E(), E(), E(), E()
    ref.record()
    torch.cuda.synchronize()
    with torch.cuda.stream(sA):
        c0.record(sA)
        if kind == "D2H":
            h.copy_(d, non_blocking=True)
        else:
            d.copy_(h, non_blocking=True)
        c1.record(sA)
    with torch.cuda.stream(sB):
        k0.record(sB)
        y = x * 2.0
        k1.record(sB)
    torch.cuda.synchronize()
    print("%s copy %.2f -> %.2f ms | kernel on the other stream %.2f -> %.2f ms" % (kind, ref.elapsed_time(c0), ref.elapsed_time(c1), ref.elapsed_time(k0), ref.elapsed_time(k1)))
