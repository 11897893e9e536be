"""Where does the double-buffered end-to-end loop of bench.py spend its time?  Same loop, CUDA events on every phase."""
import sys, os, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import gfb_loader
pkg = gfb_loader.load()
dev = pkg.GPU()
cfg = bench.workload("sqg4096")
n = 4096
prob = pkg.SurfaceQG.Problem(dev, nx=n, Lx=cfg["Lx"], nu=cfg["nu"], nnu=cfg["nnu"], H=cfg["H"], dt=cfg["dt"], stepper=cfg["stepper"])
bh0 = bench.initial_bh(n); bh0[0, 0] = 0
host_sol = torch.from_numpy(np.ascontiguousarray(bh0)).pin_memory()
diag = pkg.SurfaceQG.kinetic_energy
if os.environ.get("COMPUTE_STREAM", "0") == "1":
    torch.cuda.set_stream(torch.cuda.Stream())
main = torch.cuda.current_stream()
print("compute stream", main.cuda_stream, "max connections", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"))
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
NB = 3
bufs = [torch.empty_like(prob.sol) for _ in range(NB)]
hout = [torch.empty_like(host_sol).pin_memory() for _ in range(NB)]
E = lambda: torch.cuda.Event(enable_timing=True)
ev_in, ev_res, ev_out = ([torch.cuda.Event() for _ in range(NB)] for _ in range(3))
marks = []
MODE = os.environ.get("MODE", "step")

def upload(i):
    k = i % NB
    with torch.cuda.stream(s_in):
        s_in.wait_event(ev_out[k])
        a, z = E(), E()
        a.record(s_in)
        bufs[k].copy_(host_sol, non_blocking=True)
        z.record(s_in)
        ev_in[k].record(s_in)
    marks.append(("up%d" % i, a, z))

def run(nsteps):
    t_host = []
    upload(0)
    for i in range(nsteps):
        k = i % NB
        th0 = time.perf_counter()
        if i + 1 < nsteps:
            upload(i + 1)
        main.wait_event(ev_in[k])
        a, z = E(), E()
        a.record(main)
        prob.sol = bufs[k]
        if MODE == "dummy":
            for _ in range(40):
                bufs[k].mul_(1.0000001)
        else:
            pkg.stepforward_(prob, 1)
        z.record(main)
        ev_res[k].record(main)
        marks.append(("step%d" % i, a, z))
        th1 = time.perf_counter()
        if MODE == "step":
            diag(prob)
        else:
            main.synchronize()
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_res[k])
            a2, z2 = E(), E()
            a2.record(s_out)
            hout[k].copy_(bufs[k], non_blocking=True)
            z2.record(s_out)
            ev_out[k].record(s_out)
        marks.append(("down%d" % i, a2, z2))
        t_host.append((th1 - th0, time.perf_counter() - th1))
    torch.cuda.synchronize()
    return t_host

run(3)
marks.clear()
ref = E(); ref.record(main); torch.cuda.synchronize()
t0 = time.perf_counter()
th = run(8)
print("wall per step %.3f ms" % ((time.perf_counter() - t0) / 8 * 1e3))
for name, a, z in marks:
    print("%-7s start %7.2f  dur %6.2f ms" % (name, ref.elapsed_time(a), a.elapsed_time(z)))
print("host: enqueue ms, diag-wait ms:", [(round(x * 1e3, 2), round(y * 1e3, 2)) for x, y in th])
