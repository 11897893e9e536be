#!/usr/bin/env python
"""bench.py -- headline benchmark of the pseudospectral timestep hot path.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA library through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the CPU restatement of the
                                                           reference path (oracle/), all host threads

Metric (BASELINE.json): grid-point-steps/s = nx*ny*steps / seconds.  Workload at N = 1: configs[3],
"SurfaceQG decaying 4096^2 ETDRK4 Float64 on 1 B200 (HBM-roofline headline)", synthetic random-phase
initial buoyancy.  One "step" = one ETDRK4 step = 4 x (calcN! + stage update).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import re
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid-point-steps/sec"
UNIT = "grid-point-steps/s"
ALGO_BYTES_PER_POINT_STEP = 480.0  # 60 field passes x 8 B (Float64), RK4/ETDRK4 families
# roofline.traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from an `ncu --set full`
# capture of THIS command; it cannot be measured inside a timed run, so it is read from profiles/ncu_traffic.json, which
# tools/ncu_traffic.py writes from the .ncu-rep of the same build (keyed by kernel function; null when no entry matches)
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def kernel_functions(n, info):
    """which kernel serves each class for this context (gfb_path_info); spelled as ncu prints the instantiation, which is
    the key of profiles/ncu_traffic.json"""
    ws_r, ws_c = info.get("row_ws"), info.get("col_ws")
    if n == 16384:  # 16384-point lines: rows at half length, columns as 4 x 4096 (lane-paired K1 unless GFB_K1_16K_DIT=0)
        dit = os.environ.get("GFB_K1_16K_DIT", "1") != "0"
        lp = os.environ.get("GFB_K3_16K_LP", "0") == "1"
        return {"k_col_inv": "k_col_inv16k_dit<4096,PrepSingle>" if dit else "k_col_inv16k<4096,PrepSingle>",
                "k_row": "k_row_product_half<8192,0>", "k_col_fwd": "k_col_fwd16k_lp<4096>" if lp else "k_col_fwd16k<4096>"}
    return {"k_col_inv": ("k_col_inv_ws<%d,%d>" % (n, int(os.environ.get("GFB_K1WS_VAR", "16")))) if ws_c else "k_col_inv_single<%d>" % n,
            "k_row": ("k_row_product_ws<%d>" % n) if ws_r else (("k_row_product_tma<%d>" % n) if info.get("row_tma") else "k_row_product<%d,0>" % n),
            "k_col_fwd": ("k_col_fwd_ws<%d>" if info.get("col_ws3") else "k_col_fwd<%d,1>") % n}


def ncu_traffic(function, key="dram_bytes_per_launch"):
    try:
        return json.load(open(NCU_TRAFFIC_FILE)).get(function, {}).get(key)
    except Exception:
        return None


def workload(name):
    """sqgN: BASELINE.json configs[3] (SurfaceQG decaying, ETDRK4; examples/surfaceqg_decaying.jl scaled to N).
    twodnsN: configs[4] scaled to N (TwoDNavierStokes stochastically forced, ETDRK4;
    examples/twodnavierstokes_stochasticforcing.jl: nu = 2e-7, nnu = 2, mu = 0.1, k_f = 14, eps = 0.1 at 256^2)."""
    kind, n = ("twodns", int(name[6:])) if name.startswith("twodns") else ("sqg", int(name[3:]))
    if kind == "sqg":
        return dict(module="SurfaceQG", nx=n, ny=n, Lx=2 * np.pi, nu=1e-19 * (256 / n) ** 8, nnu=4, H=np.inf,
                    dt=0.03 * 256 / n, stepper="ETDRK4")
    return dict(module="TwoDNavierStokes", nx=n, ny=n, Lx=2 * np.pi, nu=2e-7 * (256 / n) ** 4, nnu=2, mu=0.1, nmu=0,
                dt=0.005 * 256 / n, stepper="ETDRK4", kf=14.0, dkf=1.5, eps=0.1)


def forcing_fh(cfg, seed=1):
    """One realisation of the reference example's white-in-time ring forcing, F_h = sqrt(spectrum) e^{i theta} / sqrt(dt),
    spectrum ~ exp(-(|k| - k_f)^2 / (2 dk^2)) normalised to the energy input rate eps; frozen over the timed steps."""
    n, dt = cfg["nx"], cfg["dt"]
    rng = np.random.default_rng(seed)
    kr = np.arange(n // 2 + 1)[None, :]
    l = np.fft.fftfreq(n, 1 / n)[:, None]
    K2 = (kr**2 + l**2).astype(np.float64)
    K = np.sqrt(K2)
    spec = np.exp(-(K - cfg["kf"]) ** 2 / (2 * cfg["dkf"] ** 2))
    spec[K < 2] = 0
    spec[K > 20] = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        e0 = np.sum(np.where(K2 > 0, spec / K2, 0.0)) / (n * n) ** 2 * 2  # ~ parsevalsum(spec * invKrsq / 2) per unit area
    spec *= cfg["eps"] / e0
    th = rng.uniform(0, 2 * np.pi, K.shape)
    Fh = np.sqrt(spec) * np.exp(1j * th) / np.sqrt(dt)
    Fh[0, 0] = 0
    return Fh


def initial_bh(n, seed=0):
    """Random-phase half-spectrum b_h = A(|k|) e^{i theta}, A ~ |k| (1 + (|k|/6)^4)^-1/2, rms b = 0.1, in the
    unnormalised-rfft convention; Hermitian-consistent because it is the rfft of a real field."""
    rng = np.random.default_rng(seed)
    kr = np.arange(n // 2 + 1)[None, :]
    l = np.fft.fftfreq(n, 1 / n)[:, None]
    K = np.sqrt(kr**2 + l**2)
    A = K / np.sqrt(1 + (K / 6) ** 4)
    th = rng.uniform(0, 2 * np.pi, K.shape)
    bh = A * np.exp(1j * th)
    b = np.fft.irfft2(bh, s=(n, n))
    b *= 0.1 / np.sqrt(np.mean(b * b))  # rms b = 0.1: CFL = max|u| dt/dx stays < 1 at dt = 0.03*256/n
    return np.fft.rfft2(b)


def local_spectral_columns(cfg, cols, seed, what):
    """Columns `cols` (global kr indices) of the synthetic spectra of the forced workload, generated WITHOUT ever forming
    the global array (a 16384^2 spectrum is 2 GB; every rank builds only its own columns).  what = "ic": random-phase
    state A(|k|) e^{i theta}, A ~ |k| (1 + (|k|/6)^4)^-1/2, un-normalised (the caller scales to rms 0.1 from the exact
    Parseval sum of A^2); "forcing": the ring forcing of forcing_fh (same spectrum, normalised on the kr <= 21 sub-grid
    that holds the whole ring).  Phases come from one generator per global column (seed, column), so the field does not
    depend on how the columns are dealt to ranks; the kr = 0 and kr = nx/2 columns are made Hermitian in l."""
    n = cfg["nx"]
    l = np.fft.fftfreq(n, 1 / n)[:, None]
    out = np.empty((n, len(cols)), dtype=np.complex128)
    if what == "forcing":
        krs = np.arange(0, 22)[None, :]
        K2s = krs**2 + l**2
        Ks = np.sqrt(K2s)
        specs = np.exp(-(Ks - cfg["kf"]) ** 2 / (2 * cfg["dkf"] ** 2))
        specs[(Ks < 2) | (Ks > 20)] = 0
        with np.errstate(divide="ignore", invalid="ignore"):
            e0 = np.sum(np.where(K2s > 0, specs / K2s, 0.0)) / (n * n) ** 2 * 2
    for i, c in enumerate(cols):
        K = np.sqrt(float(c) ** 2 + l[:, 0] ** 2)
        th = np.random.default_rng([seed, int(c)]).uniform(0, 2 * np.pi, n)
        if what == "ic":
            a = K / np.sqrt(1 + (K / 6) ** 4)
        else:
            a = np.exp(-(K - cfg["kf"]) ** 2 / (2 * cfg["dkf"] ** 2))
            a[(K < 2) | (K > 20)] = 0
            a = np.sqrt(a * cfg["eps"] / e0) / np.sqrt(cfg["dt"])
        col = a * np.exp(1j * th)
        if c == 0 or c == n // 2:  # a real field: a(-l) = conj a(l), a(0) and a(ny/2) real
            col[n // 2 + 1:] = np.conj(col[1:n // 2][::-1])
            col[0] = col[0].real
            col[n // 2] = col[n // 2].real
        out[:, i] = col
    if what == "ic" or True:
        z = np.nonzero(np.asarray(cols) == 0)[0]
        if len(z):
            out[0, z[0]] = 0
    return out


def ic_parseval_sumsq(n):
    """sum over the FULL spectrum of |A|^2 for the "ic" amplitude: mean(b^2) = this / n^4 (Parseval), evaluated column by
    column in float64 without a global array"""
    l2 = np.fft.fftfreq(n, 1 / n) ** 2
    tot = 0.0
    for c in range(n // 2 + 1):
        K2 = float(c) ** 2 + l2
        a2 = K2 / (1 + (K2 / 36.0) ** 2)
        if c == 0:
            a2[0] = 0
        tot += (1.0 if c in (0, n // 2) else 2.0) * float(np.sum(a2))
    return tot


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML is polled every few
    milliseconds (the timed region of the default run is a fraction of a second); nvidia-smi is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, threading.Event()
        self.sm, self.mx, self.reasons = [], [], set()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)))
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for nm, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                self.reasons.add(nm)

    def _poll_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        r = [c.strip() for c in out.strip().split(",")]
        self.sm.append(float(r[0]))
        self.mx.append(float(r[1]))
        for nm, v in zip(self.NAMES, r[3:7]):
            if v.lower().startswith("active"):
                self.reasons.add(nm)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self._poll_nvml()
                else:
                    self._poll_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.003 if self.nvml is not None else 0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def nvlink_kib(index, why=None):
    """NVLink data bytes of GPU `index` since boot, summed over its links, in KiB: (tx, rx).  NVML's
    NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX / _RX field values (payload only, no protocol overhead), all links at once (scope
    UINT_MAX) or link by link; `nvidia-smi nvlink -gt d` as the fallback.  None if nothing works (`why` collects the reasons)."""
    why = [] if why is None else why
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        fids = (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX)

        def read(scope):
            out = []
            for fid in fids:
                v = pynvml.nvmlDeviceGetFieldValues(h, [(fid, scope)])[0]
                if v.nvmlReturn != 0:
                    raise RuntimeError("field %d scope %d: nvmlReturn %d" % (fid, scope, v.nvmlReturn))
                out.append(int(v.value.ullVal))
            return out
        try:
            return tuple(read(0xFFFFFFFF))
        except Exception as e:  # noqa: BLE001
            why.append("all-links scope: %s" % e)
        tot, nlinks = [0, 0], 0
        for link in range(18):
            try:
                a = read(link)
                tot[0] += a[0]
                tot[1] += a[1]
                nlinks += 1
            except Exception as e:  # noqa: BLE001
                if link == 0:
                    why.append("per-link scope: %s" % e)
        if nlinks:
            return tuple(tot)
    except Exception as e:  # noqa: BLE001
        why.append("nvml: %s" % e)
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=10).stdout
        tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
        rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
        if tx or rx:
            return tx, rx
        why.append("nvidia-smi nvlink -gt d: no counters in %r" % out[:200])
    except Exception as e:  # noqa: BLE001
        why.append("nvidia-smi: %s" % e)
    return None


def e2e_pipelined(prob, pkg, diag, host_sol, steps):
    """steps x (pinned host -> device state, stepforward!(prob, 1), one diagnostic to the host, device -> pinned host state),
    triple-buffered: stepforward! runs IN PLACE on the buffer the state was uploaded into (the C ABI takes the state
    pointer per call), the upload of step i+1 and the download of step i-1 ride on two copy streams.  No device-to-device
    copy anywhere: cudaMemcpy D2D is served by the same copy engines as the PCIe transfers and queued behind them
    (profiles/r2l_e2e_probe_*.txt), and the diagnostic's few bytes go back BEFORE the state download is enqueued for the
    same reason.  Returns (seconds per step, last downloaded state)."""
    import torch
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    NB = 3
    bufs = [torch.empty_like(prob.sol) for _ in range(NB)]
    hout = [torch.empty_like(host_sol).pin_memory() for _ in range(NB)]
    ev_in, ev_res, ev_out = ([torch.cuda.Event() for _ in range(NB)] for _ in range(3))
    orig = prob.sol

    def upload(i):
        k = i % NB
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_out[k])  # the download of step i - 3 out of this buffer has finished (no-op at first)
            bufs[k].copy_(host_sol, non_blocking=True)
            ev_in[k].record(s_in)

    def run(nsteps):
        upload(0)
        for i in range(nsteps):
            k = i % NB
            if i + 1 < nsteps:
                upload(i + 1)
            main.wait_event(ev_in[k])
            prob.sol = bufs[k]
            pkg.stepforward_(prob, 1)
            ev_res[k].record(main)
            diag(prob)                   # device -> host scalar, synchronises the compute stream
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_res[k])
                hout[k].copy_(bufs[k], non_blocking=True)
                ev_out[k].record(s_out)
        torch.cuda.synchronize()
        return hout[(nsteps - 1) % NB]

    try:
        run(3)
        t0 = time.perf_counter()
        last = run(steps)
        return (time.perf_counter() - t0) / steps, last
    finally:
        prob.sol = orig


def oracle_problem(cfg):
    from oracle import gfo
    gfo.set_workers(os.cpu_count() or 1)
    if cfg["module"] == "SurfaceQG":
        po = gfo.surfaceqg.Problem(nx=cfg["nx"], Lx=cfg["Lx"], nu=cfg["nu"], nnu=cfg["nnu"], H=cfg["H"], dt=cfg["dt"],
                                   stepper=cfg["stepper"])
    else:
        Fh = forcing_fh(cfg)

        def calcF(F, sol, t, clock, vars, params, grid):
            F[...] = Fh
        po = gfo.twodnavierstokes.Problem(nx=cfg["nx"], Lx=cfg["Lx"], nu=cfg["nu"], nnu=cfg["nnu"], mu=cfg["mu"], nmu=cfg["nmu"],
                                          dt=cfg["dt"], stepper=cfg["stepper"], calcF=calcF, stochastic=True)
    return gfo, po


def workload_name(cfg):
    n = cfg["nx"]
    if cfg["module"] == "SurfaceQG":
        return "SurfaceQG decaying %d^2 %s Float64 (BASELINE.json configs[3])" % (n, cfg["stepper"])
    return "TwoDNavierStokes stochastically forced %d^2 %s Float64 (BASELINE.json configs[4] at the largest supported size)" % (n, cfg["stepper"])


def cpu_steps(cfg, bh0, nsteps, warmup):
    """The reference path on the host cores: numpy elementwise (serial, as Julia's broadcast) + pocketfft
    (scipy.fft, workers = all cores, the stand-in for threaded FFTW).  Returns seconds per step."""
    gfo, po = oracle_problem(cfg)
    po.sol[...] = bh0
    po.sol[0, 0] = 0
    gfo.stepforward_(po, warmup)
    t0 = time.perf_counter()
    gfo.stepforward_(po, nsteps)
    return (time.perf_counter() - t0) / nsteps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    full = workload(args.workload)
    # bounded sample of the workload: the oracle's setup alone (32-point ETD contour tables, numpy) takes minutes at
    # 16384^2, so sizes above 4096^2 are sampled on a 4096^2 sub-problem of the same module, stepper and parameters
    # scaled the same way; grid-point-steps/s is the metric, and the CPU path's rate only falls with size
    n_full = full["nx"]
    n = min(n_full, int(os.environ.get("GFB_REF_MAX_N", "4096")))
    cfg = workload(args.workload.replace(str(n_full), str(n)))
    bh0 = initial_bh(n)
    # one untimed step measures the cost, then as many of the K steps as fit ~150 s
    t0 = time.perf_counter()
    gfo, po = oracle_problem(cfg)
    po.sol[...] = bh0
    po.sol[0, 0] = 0
    gfo.stepforward_(po, 1)
    t1 = time.perf_counter() - t0
    budget = float(os.environ.get("GFB_REF_BUDGET_S", "150"))
    k = max(1, min(args.steps, int(budget / max(t1, 1e-3))))
    t0 = time.perf_counter()
    gfo.stepforward_(po, k)
    per = (time.perf_counter() - t0) / k
    val = n * n / per
    cores = os.cpu_count() or 1
    sample = "%d of %d requested ETDRK4 steps at %d^2 timed after 1 warm-up step (bounded to ~%.0f s)" % (k, args.steps, n, budget)
    if n != n_full:
        sample += "; the %d^2 workload sampled on a %d^2 sub-problem (same module, stepper, scaled parameters)" % (n_full, n)
    cfg = full
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": n_full * n_full / val * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg),
                       "impl": "oracle port of the reference CPU path: numpy broadcast + scipy.fft(pocketfft) workers=%d; "
                               "the Julia/FFTW reference itself cannot run here (no Julia, FourierFlows.jl not vendored)" % cores},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_other_config(args):
    """Device-resident lines for BASELINE.json configs[1] (slqg1024: SingleLayerQG beta-plane, finite deformation radius,
    stochastic forcing, FilteredRK4) and configs[2] (mqg2_1024 / mqg5_2048: MultiLayerQG, imposed shear, aliased_fraction 0,
    FilteredRK4): same timing protocol and per-kernel CUDA-event breakdown as the headline line, algorithmic bytes from
    SURVEY.md 8(d) (62 F; (19 + n/2) n F per stage x 4).  One GPU."""
    import torch
    import gfb_loader
    pkg = gfb_loader.load()
    dev = pkg.GPU()
    name = args.workload
    L = 2 * np.pi
    rng = np.random.default_rng(7)
    if name.startswith("slqg"):
        n = int(name[4:])
        nlay, fields_per_step = 1, 62.0
        dt = 0.05 * 128 / n
        K = np.sqrt(np.arange(n // 2 + 1)[None, :] ** 2 + np.fft.fftfreq(n, 1 / n)[:, None] ** 2)
        spec = np.exp(-(K - 14.0) ** 2 / (2 * 1.5**2))
        spec[(K < 2) | (K > 20)] = 0
        Fh = dev.array((rng.standard_normal(spec.shape) + 1j * rng.standard_normal(spec.shape)) * np.sqrt(spec * 1e-9 / dt) * n * n,
                       torch.complex128)

        def calcF(F, sol, t, clock, vars, params, grid):
            F.copy_(Fh, non_blocking=True)
        prob = pkg.SingleLayerQG.Problem(dev, nx=n, Lx=L, beta=10.0, deformation_radius=0.35, mu=0.01, dt=dt, stepper="FilteredRK4",
                                         calcF=calcF, stochastic=True)
        label = "SingleLayerQG beta-plane stochastic %d^2 FilteredRK4 Float64 (BASELINE.json configs[1])" % n
        kfun = {"k_col_inv": "k_col_inv_single<%d>" % n, "k_row": "k_row_product<%d,0>" % n, "k_col_fwd": "k_col_fwd<%d,1>" % n}
    else:
        nlay, n = int(name[3]), int(name.split("_")[1])
        fields_per_step = 4 * (19 + nlay / 2.0) * nlay
        if nlay == 2:
            H, b, U = [0.2, 0.8], [-1.0, -1.2], [1.0, 0.0]
        else:
            H, b, U = [1.0 / nlay] * nlay, None, list(np.linspace(1.0, 0.0, nlay))
        prob = pkg.MultiLayerQG.Problem(nlay, dev, nx=n, Lx=L, f0=1.0, beta=5.0, mu=5e-2, H=H, b=b, U=U, dt=2.5e-3 * 128 / n,
                                        stepper="FilteredRK4", aliased_fraction=0)
        q0 = 1e-2 * rng.standard_normal((nlay, n, n))
        pkg.MultiLayerQG.set_q_(prob, dev.array(q0))
        label = "MultiLayerQG %d-layer %d^2 FilteredRK4 Float64 (BASELINE.json configs[2])" % (nlay, n)
        kfun = {"k_col_inv": "k_col_inv<%d,PrepMulti>" % n, "k_row": "k_row_product<%d,1>" % n, "k_col_fwd": "k_col_fwd<%d,1>" % n}
    pkg.stepforward_(prob, args.warmup)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    samples = []
    nvl_why = []
    nvl0 = nvlink_kib(local, nvl_why) if world > 1 else None
    barrier()
    for _ in range(max(1, args.repeats)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        pkg.stepforward_(prob, args.steps)
        e1.record()
        torch.cuda.synchronize()
        samples.append(e0.elapsed_time(e1))
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    ms_per_step = float(np.median(samples)) / args.steps
    ksteps = min(args.steps, 20)
    prob.profile_begin()
    pkg.stepforward_(prob, ksteps)
    torch.cuda.synchronize()
    prof = prob.profile_end()
    peak, peak_src = peaks()
    algo = fields_per_step * 8.0 * n * n
    gbs = algo / (ms_per_step * 1e-3) / 1e9
    top = max(("k_col_inv", "k_row", "k_col_fwd"), key=lambda k: prof[k]["ms"])
    print(json.dumps({
        "metric": METRIC, "value": n * n / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": label, "grid": [n, n], "layers": nlay, "kernels": kfun,
                   "l2": "per-stage working set %.0f MB vs 126 MB of L2: at 1024^2 much of a stage stays L2-resident, so the 'HBM fraction' "
                         "below can exceed what DRAM counters would show (SURVEY.md 8d)" % (fields_per_step / 4 * 8.0 * n * n / 1e6),
                   "state_finite": bool(torch.isfinite(prob.sol.real).all().item())},
        "ms_samples": [x / args.steps for x in samples],
        "step_roofline": {"algorithmic_fields_per_step": fields_per_step, "algorithmic_bytes_per_step": algo, "achieved_gbs": gbs,
                          "frac_of_measured_peak": gbs / peak, "peak": peak, "peak_source": peak_src},
        "roofline": {"bound": "hbm", "kernel": top, "kernel_function": kfun[top], "peak": peak, "unit": "GB/s",
                     "avg_launch_ms": prof[top]["ms"] / max(prof[top]["launches"], 1), "achieved": None, "frac": None, "traffic": None},
        "kernels": {k: {"ms_per_step": v["ms"] / ksteps, "launches_per_step": v["launches"] / ksteps} for k, v in prof.items()},
        "gpu_launches": int(round(sum(v["launches"] for v in prof.values()) / float(ksteps) * args.steps)),
        "clocks": sampler.summary()}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cufft"])
    ap.add_argument("--workload", default=None,
                    help="sqgN | twodnsN | slqgN | mqg<layers>_N; default: sqg4096 on one GPU (BASELINE configs[3]), twodns16384 on several (configs[4])")
    ap.add_argument("--stepper", default=None, help="override the workload's stepper (e.g. FilteredRK4: the RK4 form of north_star's 60 %% target)")
    ap.add_argument("--repeats", type=int, default=5, help="the K timed steps are repeated this often; the median is reported")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cufft", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: p2p = kernels store into peer memory over NVLink (fused exchange); nccl = all-to-all")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload is None:
        args.workload = "sqg4096" if world == 1 else "twodns16384"
    if args.impl == "reference":
        return run_reference(args)
    if args.workload.startswith("slqg") or args.workload.startswith("mqg"):
        return run_other_config(args)

    import torch
    import gfb_loader
    pkg = gfb_loader.load()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    dev = pkg.GPU(local)
    torch.cuda.set_device(local)
    cfg = workload(args.workload)
    if args.stepper:  # side line: the cuFFT arm and the CPU sample restate the ETDRK4 sequence, so they are skipped
        cfg["stepper"] = args.stepper
        args.no_cufft = args.no_cpu_baseline = True
    n = cfg["nx"]
    # N > 1: ONE problem slab-sharded over the N GPUs (strong scaling): y-slabs in physical space, round-robin kr column
    # pairs in spectral space, one exchange per 2-D transform direction and stage
    shard = (rank, world, None, args.transport) if world > 1 else None
    forced = cfg["module"] != "SurfaceQG"
    t_setup = time.perf_counter()
    if not forced:
        prob = pkg.SurfaceQG.Problem(dev, nx=n, Lx=cfg["Lx"], nu=cfg["nu"], nnu=cfg["nnu"], H=cfg["H"], dt=cfg["dt"],
                                     stepper=cfg["stepper"], shard=shard)
        diag = pkg.SurfaceQG.kinetic_energy
    else:
        fh_dev = []

        def calcF(F, sol, t, clock, vars, params, grid):  # the realisation is frozen: same work per step as a fresh draw
            F.copy_(fh_dev[0], non_blocking=True)
        prob = pkg.TwoDNavierStokes.Problem(dev, nx=n, Lx=cfg["Lx"], nu=cfg["nu"], nnu=cfg["nnu"], mu=cfg["mu"], nmu=cfg["nmu"],
                                            dt=cfg["dt"], stepper=cfg["stepper"], calcF=calcF, stochastic=True, shard=shard)
        diag = pkg.TwoDNavierStokes.energy
    cols = prob.shard.cols if world > 1 else np.arange(n // 2 + 1)
    if forced:
        # column-local generators: no rank ever forms a global 16384^2 array
        fh_dev.append(dev.array(local_spectral_columns(cfg, cols, 1, "forcing"), torch.complex128))
        bh0 = local_spectral_columns(cfg, cols, 0, "ic")
        bh0 *= 0.1 * n * n / np.sqrt(ic_parseval_sumsq(n))  # rms 0.1 (exact Parseval sum of the amplitudes)
    else:
        bh0 = initial_bh(n)
        bh0[0, 0] = 0
        if world > 1:
            bh0 = prob.shard.spec_local(bh0)
    host_sol = torch.from_numpy(np.ascontiguousarray(bh0)).pin_memory()
    prob.sol.copy_(host_sol, non_blocking=True)
    setup_s = time.perf_counter() - t_setup

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.impl == "cufft":
        if world != 1:
            raise SystemExit("bench.py --impl cufft: one GPU only (the reference's device path is single-GPU)")
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from cufft_arm import CufftArm
        arm = CufftArm(prob, fh_dev[0] if forced else None)
        ms = arm.time_steps(prob.sol, args.steps, args.warmup)
        val = n * n / (ms * 1e-3)
        print(json.dumps({"impl": "cufft", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic",
                          "config": {"workload": workload_name(cfg),
                                     "impl": "the reference's GPU launch sequence restated with torch: cuFFT (torch.fft.rfft2/irfft2) + one "
                                             "unfused elementwise kernel per broadcast statement (tools/cufft_arm.py)"}}))
        return

    # ---- device-resident throughput: W warm-up steps, then K steps timed `repeats` times; median, max over ranks -----
    pkg.stepforward_(prob, args.warmup)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    samples = []
    nvl_why = []
    nvl0 = nvlink_kib(local, nvl_why) if world > 1 else None
    barrier()
    for _ in range(max(1, args.repeats)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        pkg.stepforward_(prob, args.steps)
        e1.record()
        barrier()
        samples.append(allmax(e0.elapsed_time(e1)))
    nvl1 = nvlink_kib(local) if world > 1 else None
    barrier()  # the counter read (an nvidia-smi subprocess where NVML has no counters) takes a different time on every rank
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    ms = float(np.median(samples))
    ms_per_step = ms / args.steps
    value = n * n * args.steps / (ms * 1e-3)  # one problem, however many GPUs share it
    # per-kernel device times (CUDA events around every launch) in a separate pass, so the events do not sit in the timed one
    ksteps = min(args.steps, 20)
    prob.profile_begin()
    pkg.stepforward_(prob, ksteps)
    barrier()
    prof = prob.profile_end()
    # the same K steps taken one stepforward! call at a time with a diagnostic after each (the reference's
    # stepforward!(prob, diags, 1) loop): no step may skip the aliased band of N then
    osteps = min(args.steps, 20)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(osteps):
        pkg.stepforward_(prob, 1)
        if not forced:
            diag(prob)
    e1.record()
    barrier()
    ms_observed = allmax(e0.elapsed_time(e1)) / osteps
    finite = bool(torch.isfinite(prob.sol.real).all().item())

    # ---- end to end through the public API with host buffers ----------------------------------
    # every step: state uploaded from pinned host memory, one stepforward!, state + one diagnostic back
    out_host = torch.empty_like(host_sol).pin_memory()
    h2d = host_sol.numel() * 16 * world
    d2h = (out_host.numel() * 16 + 16) * world
    for _ in range(2):
        prob.sol.copy_(host_sol, non_blocking=True)
        pkg.stepforward_(prob, 1)
        out_host.copy_(prob.sol, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        prob.sol.copy_(host_sol, non_blocking=True)
        pkg.stepforward_(prob, 1)
        out_host.copy_(prob.sol, non_blocking=True)
        ke = diag(prob)  # device->host scalar (synchronises the stream)
    barrier()
    e2e_serial_s = (time.perf_counter() - t0) / args.e2e_steps
    e2e_s, e2e_how = e2e_serial_s, "serial: upload, stepforward!(prob, 1), download, diagnostic, one after the other on one stream"
    if world == 1:
        # the same per-step work, pipelined the way a host driver would run it: the upload of step i+1 and the
        # download of step i-1 ride on copy streams while step i computes.  Checked against the serial result; any
        # problem falls back to the serial figure.
        try:
            t_pipe, last = e2e_pipelined(prob, pkg, diag, host_sol, args.e2e_steps)
            if not torch.equal(last, out_host):
                raise RuntimeError("pipelined end-to-end result differs from the serial one")
            if t_pipe < e2e_serial_s:
                e2e_s, e2e_how = t_pipe, ("triple-buffered: stepforward!(prob, 1) runs in place on the buffer the state was uploaded into; "
                                          "the pinned-host upload of step i+1 and the download of step i-1 ride on two copy streams under "
                                          "step i + its diagnostic (PCIe full duplex; no device-to-device copies)")
        except Exception as exc:  # noqa: BLE001
            sys.stderr.write("bench.py: pipelined end-to-end leg skipped: %r\n" % (exc,))
    e2e_s = allmax(e2e_s)
    e2e_val = n * n / e2e_s

    # ---- cuFFT comparison: the reference's GPU launch sequence (cuFFT + unfused broadcasts) on the same GPU, same problem
    cufft = None
    if world == 1 and not args.no_cufft and n <= 8192:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from cufft_arm import CufftArm
            arm = CufftArm(prob, fh_dev[0] if forced else None)
            cms = arm.time_steps(host_sol.to(prob.sol.device), 10, 3)
            cufft = {"ms_per_step": cms, "value": n * n / (cms * 1e-3), "unit": UNIT, "steps": 10, "warmup": 3,
                     "speedup_of_this_library": cms / ms_per_step,
                     "what": "torch.fft.rfft2/irfft2 (cuFFT) + one torch elementwise kernel per broadcast statement of the reference's "
                             "calcN!/dealias!/ETDRK4 sequence (tools/cufft_arm.py), same GPU, same workload, CUDA events"}
            del arm
            torch.cuda.empty_cache()
        except Exception as exc:  # noqa: BLE001
            sys.stderr.write("bench.py: cuFFT comparison skipped: %r\n" % (exc,))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # all-to-all volume per GPU per step: W (3 fields) + P (2 fields) buffers, (P-1)/P of each leaves the GPU
    a2a_bytes = 0.0
    if world > 1:
        a2a_bytes = 4 * 16.0 * sum(prob._a2a_elems) * (world - 1) / world
    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peak, peak_src = peaks()
    g = prob.grid
    kcut, nkr, ny = g.kr_alias0, g.nkr, g.ny
    nlive = ny - (g.l_alias[1] - g.l_alias[0])         # rows that survive the dealias mask
    band = kcut * nlive                                 # live wavenumbers of a dealiased state
    # K3 epilogue bytes per output wavenumber, ETDRK4 stages 1-4 (real coefficient tables):
    epi = [16 + 4 * 8 + 3 * 16, 16 + 3 * 8 + 16 + 2 * 16, 2 * 16 + 16 + 3 * 8 + 2 * 16, 16 + 8 + 16]

    def k3_bytes(ncols, nout):
        return 2 * 16.0 * ny * ncols + sum(epi) / 4.0 * nout
    if forced:                                          # + F_h in every stage, prevsol written in stage 1
        epi = [e + 16 for e in epi]
        epi[0] += 16
    # profiled pass: ksteps steps in one stepforward! call, the last one un-pruned; the forced problem steps one call at a
    # time (calcF! per step) and only its last step computes the aliased band
    wfull = 1.0 / float(ksteps)
    kbytes = {                                          # the pruned launches skip the aliased band of N
        "k_col_inv": (16.0 + 8.0) * band + 3 * 16.0 * ny * kcut,
        "k_row": 3 * 16.0 * ny * kcut + 2 * 16.0 * ny * (wfull * nkr + (1 - wfull) * kcut),
        "k_col_fwd": wfull * k3_bytes(nkr, nkr * ny) + (1 - wfull) * k3_bytes(kcut, band),
    }
    kbytes = {k: v / world for k, v in kbytes.items()}  # each GPU launches the kernels on its 1/N share
    info = prob.path_info()
    kfun = kernel_functions(n, info)
    top = max(("k_col_inv", "k_row", "k_col_fwd"), key=lambda k: prof[k]["ms"])
    avg_ms = prof[top]["ms"] / max(prof[top]["launches"], 1)
    achieved = kbytes[top] / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else None
    step_gbs = ALGO_BYTES_PER_POINT_STEP * n * n / world / (ms_per_step * 1e-3) / 1e9  # per GPU
    moved = sum(kbytes.values()) * 4                    # bytes the kernels of one (pruned) step must move, per GPU
    launches_per_step = sum(v["launches"] for v in prof.values()) / float(ksteps)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(cfg),
                   "grid": [n, n], "stepper": cfg["stepper"], "dt": cfg["dt"],
                   "initial_condition": "random-phase field_h ~ |k|(1+(|k|/6)^4)^-1/2, rms = 0.1"
                   + ("; ring forcing k_f = 14, one frozen realisation re-applied by calcF! every step" if forced else ", numpy default_rng(0)"),
                   "l2": "working set (>= 12 fields x %.0f MB) exceeds the 126 MB L2; no flush needed" % (16.0 * ny * nkr / 1e6),
                   "timing": "%d warm-up steps, then %d steps per stepforward! call timed %d times with CUDA events (max over ranks); median "
                             "reported, all samples in ms_samples" % (args.warmup, args.steps, len(samples)),
                   "parallelism": "1 GPU" if world == 1 else
                   "%d GPUs, one %d^2 problem: y-slabs / round-robin kr column pairs, 2 exchanges per stage via %s "
                   "(%.0f MB sent per GPU per step)" % (world, n, "P2P stores of K1/K2 into peer memory + flag barriers"
                                                        if prob.shard.transport == "p2p" else "NCCL all-to-all", a2a_bytes / 1e6),
                   "kernels": kfun, "setup_s": setup_s, "state_finite": finite},
        "ms_samples": [x / args.steps for x in samples],
        "ms_per_step_observed_every_step": ms_observed,
        "roofline": {"bound": "hbm", "kernel": top, "kernel_function": kfun[top], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None,
                     "traffic": ncu_traffic(kfun[top]) if world == 1 else None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": kbytes[top], "avg_launch_ms": avg_ms,
                     # the dominant kernel is bound by the FP64 pipe, not by HBM (five transforms per 131 KB in): its share of
                     # FP64-pipe-active cycles from the same ncu capture as `traffic` (sm__inst_executed_pipe_fp64, % of peak)
                     "fp64_pipe_frac_ncu": (ncu_traffic(kfun[top], "fp64_pipe_pct") or 0) / 100.0 or None if world == 1 else None,
                     "all_kernels": {k: {"function": kfun[k],
                                         "achieved_gbs": kbytes[k] / (prof[k]["ms"] / max(prof[k]["launches"], 1) * 1e-3) / 1e9
                                         if prof[k]["ms"] > 0 else None,
                                         "algorithmic_bytes_per_launch": kbytes[k], "traffic": ncu_traffic(kfun[k]) if world == 1 else None}
                                     for k in kbytes}},
        "step_roofline": {"algorithmic_bytes_per_step": ALGO_BYTES_PER_POINT_STEP * n * n, "achieved_gbs_per_gpu": step_gbs,
                          "frac_of_measured_peak": step_gbs / peak, "frac_of_8tbs_nominal": step_gbs / 8000.0,
                          "bytes_moved_per_pruned_step_per_gpu": moved,
                          "frac_of_measured_peak_counting_only_moved_bytes": moved / (ms_per_step * 1e-3) / 1e9 / peak,
                          "observed_every_step": {"ms_per_step": ms_observed,
                                                  "frac_of_measured_peak": ALGO_BYTES_PER_POINT_STEP * n * n / world / (ms_observed * 1e-3) / 1e9 / peak,
                                                  "what": "stepforward!(prob, 1) + diagnostic per step: every step computes the aliased band of N"}},
        "kernels": {k: {"ms_per_step": v["ms"] / ksteps, "launches_per_step": v["launches"] / ksteps} for k, v in prof.items()},
        "gpu_launches": int(round(launches_per_step * args.steps)),
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3, "serial_ms_per_step": e2e_serial_s * 1e3,
                "what": "per step: pinned-host sol -> device, stepforward!(prob, 1), sol + energy diagnostic -> host; " + e2e_how},
        "clocks": sampler.summary(),
    }
    if cufft is not None:
        line["cufft_baseline"] = cufft
    if world > 1:
        # strong-scaling reference: the SAME workload measured on one GPU of this pool (committed bench line; the 16384^2
        # problem needs 46 GB and fits one B200), so that N-GPU lines can be read as t1 / (N tN) without re-running it
        refs = {"twodns16384": "r2am_twodns16384_1gpu.json", "sqg4096": "r2x_bench100.json"}
        try:
            ref = json.loads(open(os.path.join(ROOT, "profiles", refs[args.workload])).read().strip().splitlines()[-1])
            if ref["n_gpus"] == 1 and ref["config"]["workload"] == workload_name(cfg):
                line["strong_scaling"] = {"one_gpu_ms_per_step": ref["ms_per_step"], "source": "profiles/" + refs[args.workload],
                                          "speedup": ref["ms_per_step"] / ms_per_step,
                                          "efficiency": ref["ms_per_step"] / (world * ms_per_step)}
        except Exception:  # noqa: BLE001
            pass
        # NVLink roofline of the sharded step: bytes that must leave each GPU / measured peer bandwidth (770 GB/s/dir)
        line["nvlink"] = {"a2a_bytes_per_gpu_per_step": a2a_bytes, "floor_ms_at_770GBs": a2a_bytes / 770e9 * 1e3,
                          "frac_of_nvlink_floor": (a2a_bytes / 770e9 * 1e3) / ms_per_step,
                          "exchanges_per_step": 8, "transport": prob.shard.transport}
        if nvl0 is not None and nvl1 is not None:
            # measured on rank 0's GPU over the timed repeats (NVML link counters, payload bytes)
            nst = args.steps * max(1, args.repeats)
            tx, rx = (nvl1[0] - nvl0[0]) * 1024.0 / nst, (nvl1[1] - nvl0[1]) * 1024.0 / nst
            line["nvlink"].update({"measured_tx_bytes_per_step": tx, "measured_rx_bytes_per_step": rx,
                                   "measured_tx_gbs": tx / (ms_per_step * 1e-3) / 1e9,
                                   "measured_vs_model": tx / a2a_bytes if a2a_bytes else None,
                                   "measured_how": "NVML NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX/RX deltas of this rank's GPU over the timed repeats"})
        elif nvl_why:
            line["nvlink"]["measured_unavailable"] = "; ".join(nvl_why)[:400]
    if world == 1 and not args.no_cpu_baseline and n <= 4096:  # larger grids: `--impl reference` samples them at 4096^2
        nsamp = 2 if n >= 4096 else 4
        full0 = bh0 if not forced else None
        per = cpu_steps(cfg, full0, nsamp, 1) if full0 is not None else cpu_steps(cfg, np.ascontiguousarray(bh0), nsamp, 1)
        line["cpu_baseline"] = {"value": n * n / per, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": "%d ETDRK4 steps at %d^2 after 1 warm-up step; numpy broadcast + scipy.fft workers=%d"
                                          % (nsamp, n, os.cpu_count() or 1), "ms_per_step": per * 1e3}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
