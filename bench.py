#!/usr/bin/env python
"""bench.py — headline benchmark of the nd4js batched dense-LA hot path on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line.

Metric (BASELINE.json): matrices/s (+ fp64 GFLOP/s, % roofline) of the broadcast-batched
nd.la.matmul float64 [65536,32,32] x [65536,32,32] (configs[1], "C2"); one step = one pass of the
hot path over one batch.  Weak scaling: every rank (one process per GPU) owns its own batch of
65 536 matrix pairs, no data-path collective.

  value     matrices/s with inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e       the same metric through the host-buffer C ABI (nd4b_matmul_f64, what nd.la.matmul2 calls):
            pinned host buffers, H2D + kernels + D2H inside the timed region
  roofline  dominant kernel matmul32_kernel: 24 576 algorithmic bytes per matrix / launch time vs measured HBM peak
  cpu_baseline  the oracle (C restatement of the reference's JS loops) on one host core (the reference is
            single-threaded), on a bounded sample, rank 0 / N=1 only
`--workload c1|c3|c4|c5` times the other BASELINE configs the same way (s3, l4, i4: the 'next' rows; g4k: compute-bound probe); `others` in the default line
carries their kernel-only numbers.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, units per rank, algorithmic bytes per unit, flop per unit)
    "c1": ("nd.la.matmul float64 512x512 . 512x512 single matrix", 1, 3 * 512 * 512 * 8, 2 * 512 ** 3),
    "c2": ("broadcast-batched nd.la.matmul float64 [65536,32,32]x[65536,32,32]", 65536, 24576, 65536),
    "c3": ("batched nd.la.cholesky_decomp float64 SPD [262144,16,16]", 262144, 4096, 16 ** 3 / 3.0),
    "c4": ("batched nd.la.qr_decomp float64 [65536,64,32] Householder", 65536, 40960, 2 * (2 * 64 * 32 * 32 - 2 * 32 ** 3 / 3.0)),
    "c5": ("batched nd.la.svd_jac_1sided float64 [16384,64,64]", 16384, 98816, 2016 * 1152),  # flop per sweep
    # next row 8f-1: the solve that follows C3 (one right-hand side per matrix); bytes: L read whole + y in + x out
    "s3": ("batched nd.la.cholesky_solve float64 L[262144,16,16], y[262144,16,1]", 262144, 2048 + 128 + 128, 2 * 16 * 16),
    # next row 8f-1: least squares from C4's factors; bytes: Q, R, y in, x out
    "l4": ("batched nd.la.qr_lstsq float64 Q[65536,64,32], R[65536,32,32], y[65536,64,1]", 65536, (2048 + 1024 + 64 + 32) * 8, 2 * 64 * 32 + 32 * 32),
    # next row 8f-3: R and Q^T y without forming Q; bytes: A, y in; R (M x N), Q^T y out; flop: Householder R phase + reflectors on y
    "i4": ("batched nd.la._qr_decomp_inplace float64 A[65536,64,32], y[65536,64,1]", 65536, (2048 + 64 + 2048 + 64) * 8,
           2 * 64 * 32 * 32 - 2 * 32 ** 3 / 3.0 + 4 * 64 * 32),
    # compute-bound probe of the same matmul kernel family (north_star: matmul vs FP64 tensor-core peak)
    "g4k": ("nd.la.matmul float64 4096x4096 . 4096x4096 single matrix (compute-bound probe)", 1, 3 * 4096 * 4096 * 8, 2 * 4096 ** 3),
}


# FP64 roofline denominator: the DMMA.8x8x4 peak measured on this pool's B200 (profiles/r01_fp64_peak.json; the DFMA peak
# is 36.83 — both instruction kinds share one pipe).  MEASURED_PEAKS.json holds no FP64 entry.
FP64_PEAK_TFLOPS = 37.15

# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the `ncu --set full` captures
# summarised under profiles/ (file named per entry); None where no capture of the current kernel exists.
NCU_TRAFFIC_BYTES = {
    "c2": (1.577e9, "profiles/r01_ncu_c2_final.txt"),
    "c3": (1.014e9, "profiles/r01_ncu_c3_v3.txt"),
    "c4": (2.800e9, "profiles/r01_ncu_c4_blocked.txt"),
    "c5": (1.576e9, "profiles/r01_ncu_c5_v3.txt"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """--impl reference: the reference's own algorithm on the host CPU.  No JS engine exists in this image
    (SURVEY fact 5), so the stand-in is oracle/libnd4ref.so — a line-faithful C restatement of
    nd4js's loops (kind "port"), single-threaded because the reference is."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import nd4ref
    nd4ref.build()
    desc, units, bpu, fpu = WORKLOADS[args.workload]
    rng = np.random.default_rng(3)
    sample = {"c1": 1, "c2": 8192, "c3": 32768, "c4": 2048, "c5": 24, "g4k": 1, "s3": 65536, "l4": 8192, "i4": 2048}[args.workload]
    fn, data = _ref_case(args.workload, sample, rng, nd4ref)
    for _ in range(min(args.warmup, 1)):
        fn(*data)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn(*data)
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "matrices/s", "value": val, "unit": "matrices/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "units_per_step": sample, "l2": "n/a (cpu)"},
        "gflops": val * fpu / 1e9,
        "cpu_baseline": {"value": val, "unit": "matrices/s", "cores": 1, "kind": "port",
                         "sample": "%d of %d units per step, oracle/libnd4ref.so (C restatement of the JS loops), "
                                   "host has %d cores, reference is single-threaded" % (sample, units, os.cpu_count())},
        "e2e": {"value": val, "unit": "matrices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def _ref_case(name, n, rng, nd4ref):
    import numpy as np
    if name == "c1":
        return nd4ref.matmul2, (rng.uniform(-1, 1, (512, 512)), rng.uniform(-1, 1, (512, 512)))
    if name == "g4k":
        return nd4ref.matmul2, (rng.uniform(-1, 1, (4096, 4096)), rng.uniform(-1, 1, (4096, 4096)))
    if name == "c2":
        return nd4ref.matmul2, (rng.uniform(-1, 1, (n, 32, 32)), rng.uniform(-1, 1, (n, 32, 32)))
    if name == "c3":
        g = rng.uniform(-1, 1, (n, 16, 16))
        return nd4ref.cholesky_decomp, (g @ g.transpose(0, 2, 1) + 16 * np.eye(16),)
    if name == "c4":
        return nd4ref.qr_decomp, (rng.uniform(-1, 1, (n, 64, 32)),)
    if name == "l4":
        q, r = np.linalg.qr(rng.uniform(-1, 1, (n, 64, 32)))
        return nd4ref.qr_lstsq, (q, r, rng.uniform(-1, 1, (n, 64, 1)))
    if name == "i4":
        return nd4ref.qr_decomp_inplace, (rng.uniform(-1, 1, (n, 64, 32)), rng.uniform(-1, 1, (n, 64, 1)))
    if name == "s3":
        g = rng.uniform(-1, 1, (n, 16, 16))
        return nd4ref.cholesky_solve, (np.linalg.cholesky(g @ g.transpose(0, 2, 1) + 16 * np.eye(16)), rng.uniform(-1, 1, (n, 16, 1)))
    return nd4ref.svd_jac_2sided, (rng.uniform(-1, 1, (n, 64, 64)),)


class DeviceCase:
    """One workload with device-resident inputs; step() enqueues exactly one pass on the given stream."""

    def __init__(self, name, torch, lib, dev, units):
        import numpy as np
        self.name, self.torch, self.lib, self.dev, self.units = name, torch, lib, dev, units
        g = torch.Generator(device="cuda").manual_seed(1234 + dev)
        f64 = dict(dtype=torch.float64, device="cuda")
        u = lambda *s: torch.rand(*s, generator=g, **f64) * 2 - 1
        self.sweeps = None
        self.work = None
        if name in ("c1", "g4k"):
            n = 512 if name == "c1" else 4096
            self.n = n
            self.a, self.b, self.out = u(n, n), u(n, n), [torch.empty(n, n, **f64)]
        elif name == "c2":
            self.a, self.b, self.out = u(units, 32, 32), u(units, 32, 32), [torch.empty(units, 32, 32, **f64)]
        elif name == "c3":
            gmat = u(units, 16, 16)
            self.a = torch.baddbmm(16.0 * torch.eye(16, **f64).expand(units, 16, 16), gmat, gmat.transpose(1, 2))
            del gmat
            self.out = [torch.empty(units, 16, 16, **f64)]
            self.info = torch.full((1,), 2 ** 62, dtype=torch.int64, device="cuda")
        elif name == "s3":
            gmat = u(units, 16, 16)
            self.a = torch.linalg.cholesky(torch.baddbmm(16.0 * torch.eye(16, **f64).expand(units, 16, 16), gmat, gmat.transpose(1, 2)))
            del gmat
            self.b = u(units, 16, 1)
            self.out = [torch.empty(units, 16, 1, **f64)]
        elif name == "l4":
            self.a, self.r = torch.linalg.qr(u(units, 64, 32))
            self.a, self.r = self.a.contiguous(), self.r.contiguous()
            self.b = u(units, 64, 1)
            self.out = [torch.empty(units, 32, 1, **f64)]
        elif name == "i4":
            self.a, self.b = u(units, 64, 32), u(units, 64, 1)
            self.out = [torch.empty(units, 64, 32, **f64), torch.empty(units, 64, 1, **f64)]
        elif name == "c4":
            self.a = u(units, 64, 32)
            self.out = [torch.empty(units, 64, 32, **f64), torch.empty(units, 32, 32, **f64)]
        elif name == "c5":
            self.a = u(units, 64, 64)
            self.out = [torch.empty(units, 64, 64, **f64), torch.empty(units, 64, **f64), torch.empty(units, 64, 64, **f64)]
            self.sweeps = torch.zeros(4, dtype=torch.int32, device="cuda")
            # per-matrix sweep counts of every launch are summed here: work per matrix is proportional to its sweeps
            self.sweep_sum = torch.zeros(1, dtype=torch.int64, device="cuda")
            self.launches = 0
            if lib.nd4b_dev_svd_sweep_counter(dev, C.c_void_p(self.sweep_sum.data_ptr())):
                raise RuntimeError(lib.nd4b_last_error().decode())
        torch.cuda.synchronize()

    def mean_sweeps(self):
        """Mean Jacobi sweeps per matrix over all launches so far (None for the other workloads)."""
        if self.sweeps is None or not self.launches:
            return None
        return float(self.sweep_sum.item()) / (self.launches * self.units)

    def close(self):
        if self.sweeps is not None:
            self.lib.nd4b_dev_svd_sweep_counter(self.dev, None)

    def step(self, stream):
        L, p, d = self.lib, (lambda t: C.c_void_p(t.data_ptr())), self.dev
        s = C.c_void_p(stream)
        if self.name in ("c1", "g4k"):
            rc = L.nd4b_dev_matmul_f64(d, s, p(self.a), 0, p(self.b), 0, p(self.out[0]), 1, self.n, self.n, self.n)
        elif self.name == "c2":
            rc = L.nd4b_dev_matmul_f64(d, s, p(self.a), 1024, p(self.b), 1024, p(self.out[0]), self.units, 32, 32, 32)
        elif self.name == "c3":
            rc = L.nd4b_dev_cholesky_f64(d, s, p(self.a), p(self.out[0]), self.units, 16, p(self.info))
        elif self.name == "s3":
            rc = L.nd4b_dev_tri_solve_f64(d, s, 2, p(self.a), 256, p(self.b), 16, p(self.out[0]), self.units, 16, 1)
        elif self.name == "l4":
            rc = L.nd4b_dev_qr_lstsq_f64(d, s, p(self.a), p(self.r), p(self.b), p(self.out[0]), self.units, 64, 32, 32, 1)
        elif self.name == "i4":
            rc = L.nd4b_dev_qr_inplace_f64(d, s, p(self.a), p(self.b), p(self.out[0]), p(self.out[1]), self.units, 64, 32, 1)
        elif self.name == "c4":
            rc = L.nd4b_dev_qr_f64(d, s, p(self.a), p(self.out[0]), p(self.out[1]), self.units, 64, 32, None, 0)
        else:
            rc = L.nd4b_dev_svd_jac1_f64(d, s, p(self.a), p(self.out[0]), p(self.out[1]), p(self.out[2]),
                                         self.units, 64, 64, p(self.sweeps), None, 0)
            self.launches += 1
        if rc:
            raise RuntimeError("nd4b call failed: %s" % self.lib.nd4b_last_error().decode())


def time_device(torch, case, steps, warmup):
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(warmup):
        case.step(stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        case.step(stream)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3  # seconds for `steps` launches


def host_case(name, nd, units):
    """Pinned host buffers + the public host-buffer call for the e2e leg."""
    import numpy as np
    import torch
    rng = np.random.default_rng(99)

    def pinned(shape, fill=None):
        t = torch.empty(shape, dtype=torch.float64).pin_memory()
        if fill is not None:
            t.numpy()[...] = fill
        return t

    L = nd.load()
    p = lambda t: C.c_void_p(t.data_ptr())
    if name in ("c1", "c2", "g4k"):
        sh = (512, 512) if name == "c1" else (4096, 4096) if name == "g4k" else (units, 32, 32)
        a, b, c = pinned(sh, rng.uniform(-1, 1, sh)), pinned(sh, rng.uniform(-1, 1, sh)), pinned(sh)
        shp = np.asarray(sh, np.int32)
        sp = C.c_void_p(shp.ctypes.data)
        call = lambda: L.nd4b_matmul_f64(p(a), sp, len(sh), p(b), sp, len(sh), p(c), sp, len(sh))
        return call, a.numel() * 16, c.numel() * 8, (a, b, c, shp)
    if name == "c3":
        g = rng.uniform(-1, 1, (units, 16, 16))
        s, l = pinned((units, 16, 16), g @ g.transpose(0, 2, 1) + 16 * np.eye(16)), pinned((units, 16, 16))
        bad = C.c_int64(0)
        return (lambda: L.nd4b_cholesky_f64(p(s), p(l), units, 16, C.byref(bad))), s.numel() * 8, l.numel() * 8, (s, l)
    if name == "l4":
        qq, rr = np.linalg.qr(rng.uniform(-1, 1, (units, 64, 32)))
        q, r = pinned((units, 64, 32), qq), pinned((units, 32, 32), rr)
        y, x = pinned((units, 64, 1), rng.uniform(-1, 1, (units, 64, 1))), pinned((units, 32, 1))
        return ((lambda: L.nd4b_qr_lstsq_f64(p(q), p(r), p(y), p(x), units, 64, 32, 32, 1)), (q.numel() + r.numel() + y.numel()) * 8,
                x.numel() * 8, (q, r, y, x))
    if name == "s3":
        g = rng.uniform(-1, 1, (units, 16, 16))
        l = pinned((units, 16, 16), np.linalg.cholesky(g @ g.transpose(0, 2, 1) + 16 * np.eye(16)))
        y, x = pinned((units, 16, 1), rng.uniform(-1, 1, (units, 16, 1))), pinned((units, 16, 1))
        ls, ys = np.asarray((units, 16, 16), np.int32), np.asarray((units, 16, 1), np.int32)
        lp, yp = C.c_void_p(ls.ctypes.data), C.c_void_p(ys.ctypes.data)
        return ((lambda: L.nd4b_tri_solve_f64(2, p(l), lp, 3, p(y), yp, 3, p(x), yp, 3)), (l.numel() + y.numel()) * 8, x.numel() * 8,
                (l, y, x, ls, ys))
    if name == "i4":
        a, y = pinned((units, 64, 32), rng.uniform(-1, 1, (units, 64, 32))), pinned((units, 64, 1), rng.uniform(-1, 1, (units, 64, 1)))
        r, qty = pinned((units, 64, 32)), pinned((units, 64, 1))
        return ((lambda: L.nd4b_qr_inplace_f64(p(a), p(y), p(r), p(qty), units, 64, 32, 1)), (a.numel() + y.numel()) * 8,
                (r.numel() + qty.numel()) * 8, (a, y, r, qty))
    if name == "c4":
        a, q, r = pinned((units, 64, 32), rng.uniform(-1, 1, (units, 64, 32))), pinned((units, 64, 32)), pinned((units, 32, 32))
        return (lambda: L.nd4b_qr_f64(p(a), p(q), p(r), units, 64, 32)), a.numel() * 8, (q.numel() + r.numel()) * 8, (a, q, r)
    a = pinned((units, 64, 64), rng.uniform(-1, 1, (units, 64, 64)))
    u, sv, v = pinned((units, 64, 64)), pinned((units, 64)), pinned((units, 64, 64))
    sw = C.c_int(0)
    return ((lambda: L.nd4b_svd_jac1_f64(p(a), p(u), p(sv), p(v), units, 64, 64, C.byref(sw))), a.numel() * 8,
            (u.numel() + sv.numel() + v.numel()) * 8, (a, u, sv, v))


def fp64_peaks(lib, dev):
    """DFMA / DMMA pipe peaks measured on this GPU now (MEASURED_PEAKS.json has no fp64 entry)."""
    out = {}
    ms = C.c_float(0)
    iters, blocks, threads = 4096, 148 * 8, 256
    if lib.nd4b_probe_fp64(dev, 0, iters, blocks, threads, C.byref(ms)) == 0 and ms.value > 0:
        out["dfma_tflops"] = 2.0 * 16 * iters * blocks * threads / (ms.value * 1e-3) / 1e12
    if lib.nd4b_probe_fp64(dev, 1, iters, blocks, threads, C.byref(ms)) == 0 and ms.value > 0:
        out["dmma_tflops"] = 512.0 * 8 * iters * blocks * (threads // 32) / (ms.value * 1e-3) / 1e12
    return out


def pin_to_gpu_numa_node(index):
    """One process per GPU: run this rank (and the library's copy threads it creates, and the first touch of its pinned
    staging buffers) on the CPUs next to its GPU, so that the eight PCIe streams of a box do not all cross the socket
    interconnect.  Best effort (NVML's ideal CPU set for the device); returns the number of CPUs or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import nd4js_b200 as nd
    from nd4js_b200.partition import max_over_ranks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = pin_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its banner / debug lines to stdout by default; stdout carries ONE JSON line, so send them to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nd.init([local])
    lib = nd.load()
    desc, units, bpu, fpu = WORKLOADS[args.workload]
    hbm_peak, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident (value + roofline) ----------------
    case = DeviceCase(args.workload, torch, lib, local, units)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    secs = time_device(torch, case, args.steps, max(args.warmup, 3))
    barrier()
    secs = max_over_ranks(secs, device="cuda")
    sweeps = int(case.sweeps[0].item()) if case.sweeps is not None else None
    sweeps_mean = case.mean_sweeps()
    value = world * units * args.steps / secs
    launch_s = secs / args.steps
    flop_unit = fpu * (sweeps_mean if sweeps_mean else 1)  # Jacobi work is proportional to the sweeps each matrix needed
    achieved_gbs = bpu * units / launch_s / 1e9
    case.close()
    del case
    torch.cuda.empty_cache()

    if args.kernel_only:  # tuning aid: device-resident timing only, not a bench line
        if sampler:
            sampler.stop()
        if rank == 0:
            emit({"workload": args.workload, "ms_per_launch": 1e3 * launch_s, "matrices_per_s": value,
                  "gflops": value * flop_unit / 1e9, "hbm_gbs_algorithmic": achieved_gbs, "sweeps_max": sweeps,
                  "sweeps_mean": sweeps_mean})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- end to end through the host-buffer C ABI ----------------
    call, h2d, d2h, keep = host_case(args.workload, nd, units)
    e2e_steps = max(1, min(args.steps, 10))
    for _ in range(2):
        if call():
            raise RuntimeError(lib.nd4b_last_error().decode())
    s0 = nd.stats()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        if call():
            raise RuntimeError(lib.nd4b_last_error().decode())
    torch.cuda.synchronize()
    e2e_secs = time.perf_counter() - t0
    barrier()
    e2e_secs = max_over_ranks(e2e_secs, device="cuda")
    clocks = sampler.stop() if sampler else None  # sampled over both timed regions (device-resident + e2e)
    s1 = nd.stats()
    e2e_value = world * units * e2e_steps / e2e_secs
    e2e_launches = s1["kernel_launches"] - s0["kernel_launches"]
    del keep

    line = None
    if rank == 0:
        kernel = {"c1": "gemm_pipe_kernel", "g4k": "gemm_pipe_kernel", "c2": "matmul32_kernel", "c3": "chol16_kernel",
                  "c4": "qr64x32_blocked_kernel", "c5": "svd64cb_kernel", "s3": "trisolve16_kernel", "l4": "qr_lstsq32_kernel", "i4": "qr64x32_inplace_kernel"}[args.workload]
        traffic, traffic_src = NCU_TRAFFIC_BYTES.get(args.workload, (None, None))
        if args.workload in ("c1", "g4k", "c5"):
            # compute-bound configs: against the FP64 pipe — DMMA for the GEMMs, DFMA for the Jacobi SVD; both
            # instruction kinds share one pipe and one peak on B200
            tflops = value / world * flop_unit / 1e12
            roofline = {"bound": "tensor", "achieved": tflops, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                        "frac": tflops / FP64_PEAK_TFLOPS, "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": "measured FP64 pipe peak (profiles/r01_fp64_peak.json: DMMA.8x8x4 37.15, DFMA 36.83 TFLOP/s)",
                        "kernel": kernel, "algorithmic_flop_per_unit": flop_unit, "units_per_launch": units}
        else:
            roofline = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved_gbs / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": peak_src, "kernel": kernel, "algorithmic_bytes_per_unit": bpu, "units_per_launch": units}
        line = {
            "metric": "matrices/s", "value": value, "unit": "matrices/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * launch_s, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "units_per_gpu": units, "inputs": "U(-1,1) (SPD G.G^T+16I for cholesky), seeded",
                       "l2": "inputs+outputs per step (%.0f MiB) exceed the 126 MB L2; no explicit flush" % (bpu * units / 2 ** 20)
                       if bpu * units > 200e6 else "working set fits L2: kernel-only figure is L2-warm, see e2e"},
            "gflops": value * flop_unit / 1e9,
            "e2e": {"value": e2e_value, "unit": "matrices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "timing": "host wall clock around the blocking C-ABI call, barrier+sync both sides, max over ranks",
                    "api": "nd4b_%s_f64 (host buffers, pinned)" % {"c1": "matmul", "c2": "matmul", "g4k": "matmul", "c3": "cholesky", "c4": "qr", "c5": "svd_jac1", "s3": "tri_solve", "l4": "qr_lstsq", "i4": "qr_inplace"}[args.workload],
                    "rank_cpus_near_gpu": numa},
            "gpu_launches": args.steps + int(e2e_launches),
            "roofline": roofline,
            "clocks": clocks,
        }
        if sweeps:
            line["sweeps_max"], line["sweeps_mean"] = sweeps, sweeps_mean
        if args.fp64_probes:
            line["fp64_peaks_measured"] = fp64_peaks(lib, local)
        else:
            try:
                line["fp64_peaks_measured"] = {k: v for k, v in json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json"))).items()
                                               if k.endswith("_peak_tflops") or k == "copy_gbs"}
                line["fp64_peaks_measured"]["source"] = "profiles/r01_fp64_peak.json (tools/fp64_peak.py on this pool's B200)"
            except Exception:
                pass

    # ---------------- the other BASELINE configs, kernel-only, short ----------------
    if args.workload == "c2" and not args.no_others:
        others = {}
        for name in ("c1", "g4k", "c3", "s3", "c4", "l4", "i4", "c5"):
            d2, u2, b2, f2 = WORKLOADS[name]
            c2 = DeviceCase(name, torch, lib, local, u2)
            barrier()
            n2 = 3 if name == "c5" else 10
            s2 = time_device(torch, c2, n2, 3)
            s2 = max_over_ranks(s2, device="cuda")
            sw = c2.mean_sweeps() or 1
            gf = world * u2 * n2 / s2 * f2 * sw / 1e9
            gbs = b2 * u2 * n2 / s2 / 1e9
            others[name] = {"workload": d2, "matrices_per_s": world * u2 * n2 / s2, "ms_per_launch": 1e3 * s2 / n2,
                            "gflops": gf, "hbm_gbs_algorithmic": gbs,
                            "frac_of_hbm_peak": gbs / world / hbm_peak, "frac_of_fp64_peak": gf / world / 1e3 / FP64_PEAK_TFLOPS,
                            "sweeps_mean": sw if name == "c5" else None,
                            "sweeps_max": int(c2.sweeps[0].item()) if name == "c5" else None}
            c2.close()
            del c2
            torch.cuda.empty_cache()
        if rank == 0:
            line["others"] = others

    # ---------------- CPU baseline: rank 0, N=1 only ----------------
    if rank == 0 and world == 1 and not args.no_cpu:
        import numpy as np
        from oracle import nd4ref
        nd4ref.build()
        sample = {"c1": 1, "c2": 65536, "c3": 262144, "c4": 8192, "c5": 96, "g4k": 1, "s3": 262144, "l4": 16384, "i4": 8192}[args.workload]
        fn, data = _ref_case(args.workload, sample, np.random.default_rng(3), nd4ref)
        reps, t0 = 0, time.perf_counter()
        while True:
            fn(*data)
            reps += 1
            if time.perf_counter() - t0 > 10.0 or reps >= 50:
                break
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sample * reps / dt, "unit": "matrices/s", "cores": 1, "kind": "port",
                                "sample": "%d units x %d passes (%.1f s) with oracle/libnd4ref.so, a C restatement of the JS loops "
                                          "(no JS engine in this image); host has %d cores, the reference is single-threaded"
                                          % (sample, reps, dt, os.cpu_count())}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints was diverted to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    # NCCL / torch print banners (e.g. "NCCL version ...") on fd 1: keep the contract's stdout clean
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-others", action="store_true", help="skip the kernel-only lines of the other configs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--kernel-only", action="store_true", help="tuning aid: print the device-resident timing only")
    ap.add_argument("--fp64-probes", action="store_true", help="also measure the DFMA/DMMA pipe peaks (extra launches before the timed region)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
