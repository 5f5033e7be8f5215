#!/usr/bin/env python
"""bench.py — headline benchmark of the nd4js batched dense-LA hot path on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line.

Metric (BASELINE.json): matrices/s (+ fp64 GFLOP/s, % roofline) of the broadcast-batched
nd.la.matmul float64 [65536,32,32] x [65536,32,32] (configs[1], "C2"); one step = one pass of the
hot path over one batch.  Weak scaling: every rank (one process per GPU) owns its own batch of
65 536 matrix pairs, no data-path collective.

  value     matrices/s with inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e       the same metric through the operator a user calls (nd4js_b200.la.matmul2 -> nd4b_matmul_f64): NDArrays in
            page-locked host memory in and out, H2D + kernels + D2H inside the timed region; `e2e_pageable`: the same call
            on ordinary numpy inputs (staged through the library's pinned ring)
  parity    after the timed regions a sample of every timed workload's output (first / middle / last units) is checked
            against the CPU oracle with the bars of oracle/parity.py; no value is printed when a check fails
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
    # the broadcast variant of C2: one B for all products; B is read once, so a unit moves A in and C out
    "c2b": ("broadcast-batched nd.la.matmul float64 [65536,32,32]x[1,32,32]", 65536, 16384, 65536),
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
    # next row 8f-2: the solve that follows C5; bytes: U, sv, V, y in, x out
    "v5": ("batched nd.la.svd_lstsq float64 U[16384,64,64], sv[16384,64], V[16384,64,64], y[16384,64,1]", 16384,
           (4096 + 64 + 4096 + 64 + 64) * 8, 2 * 64 * 64 * 2 + 64),
    # compute-bound probe of the same matmul kernel family (north_star: matmul vs FP64 tensor-core peak)
    "g4k": ("nd.la.matmul float64 4096x4096 . 4096x4096 single matrix (compute-bound probe)", 1, 3 * 4096 * 4096 * 8, 2 * 4096 ** 3),
}


# FP64 roofline denominator: the DMMA.8x8x4 peak measured on this pool's B200 (profiles/r01_fp64_peak.json; the DFMA peak
# is 36.83 — both instruction kinds share one pipe).  MEASURED_PEAKS.json holds no FP64 entry.
FP64_PEAK_TFLOPS = 37.15

# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the `ncu --set full` captures
# summarised under profiles/ (file named per entry); None where no capture of the current kernel exists.
NCU_TRAFFIC_BYTES = {
    "c1": (4.21e6, "profiles/r02_ncu_c1.txt"),    # the operands stay in L2 between launches: the write-back of C is not seen by DRAM
    "c2": (1.578e9, "profiles/r02_ncu_c2.txt"),
    "c3": (1.015e9, "profiles/r02_ncu_c3.txt"),
    "g4k": (1.225e9, "profiles/r02_ncu_g4k_bulk.txt"),
    "c4": (3.016e9, "profiles/r02_ncu_c4_tmap.txt"),   # 2.68e9 algorithmic + the idle column blocks parked in local memory around the panels
    # three launches per call: FP32 Jacobi 0.777 GB, hand-over 1.825 GB, FP64 Jacobi 2.154 GB (V0, G1, V1 pass through HBM)
    "c5": (4.756e9, "profiles/r02_ncu_c5_pipeline_v2.txt"),
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


KERNEL = {"c1": "gemm_pipe_kernel", "g4k": "gemm_pipe_kernel", "c2": "matmul32_kernel", "c2b": "matmul32_kernel", "c3": "chol16_kernel",
          "c4": "qr64x32_blocked_kernel", "c5": "svd64_pre32_kernel + svd64_ortho_kernel + svd64cb_kernel", "s3": "trisolve16_kernel", "l4": "qr_lstsq32_kernel",
          "i4": "qr64x32_inplace_kernel", "v5": "svd_lstsq_kernel"}
OPERATOR = {"c1": "matmul2", "c2": "matmul2", "c2b": "matmul2", "g4k": "matmul2", "c3": "cholesky_decomp", "c4": "qr_decomp",
            "c5": "svd_jac_1sided", "s3": "cholesky_solve", "l4": "qr_lstsq", "i4": "_qr_decomp_inplace", "v5": "svd_lstsq"}
# units the --impl reference arm (and the in-line cpu_baseline) time per step: the whole workload where one pass of the
# one-core port takes seconds, a stated sample where it would take minutes
REF_SAMPLE = {"c1": 1, "c2": 65536, "c2b": 65536, "c3": 262144, "c4": 8192, "c5": 96, "g4k": 1, "s3": 262144, "l4": 16384, "i4": 8192,
              "v5": 16384}


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
    sample = REF_SAMPLE[args.workload]
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
        "config": {"workload": desc, "units_per_gpu": units, "units_per_step": sample, "l2": "n/a (cpu)"},
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
    if name == "c2b":
        return nd4ref.matmul2, (rng.uniform(-1, 1, (n, 32, 32)), rng.uniform(-1, 1, (1, 32, 32)))
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
    if name == "v5":
        u, sv, v = np.linalg.svd(rng.uniform(-1, 1, (n, 64, 64)))
        return nd4ref.svd_lstsq, (u, sv, v, rng.uniform(-1, 1, (n, 64, 1)))
    return nd4ref.svd_jac_2sided, (rng.uniform(-1, 1, (n, 64, 64)),)


def sample_index(units, per=16):
    """Units checked against the oracle after a timed region: the first, the middle and the last `per` of the batch
    (first / middle / last CTAs of the launch)."""
    import numpy as np
    if units <= 3 * per:
        return np.arange(units)
    mid = units // 2
    return np.concatenate([np.arange(per), np.arange(mid - per // 2, mid - per // 2 + per), np.arange(units - per, units)])


def parity_of(name, ins, outs):
    """(max_err, bar) of one workload's sampled inputs / outputs (numpy) against the CPU oracle (oracle/parity.py)."""
    from oracle import parity
    op = {"c1": "matmul", "g4k": "matmul", "c2": "matmul", "c2b": "matmul", "c3": "cholesky", "s3": "cholesky_solve", "l4": "qr_lstsq",
          "i4": "qr_inplace", "c4": "qr", "c5": "svd", "v5": "svd_lstsq"}[name]
    err, bar = parity.check(op, *ins, *outs)
    return {"max_err": err, "bar": bar, "ok": bool(err <= bar), "op": op}


class DeviceCase:
    """One workload with device-resident inputs; step() enqueues exactly one pass on the given stream."""

    def __init__(self, name, torch, lib, dev, units):
        self.name, self.torch, self.lib, self.dev, self.units = name, torch, lib, dev, units
        g = torch.Generator(device="cuda").manual_seed(1234 + dev)
        f64 = dict(dtype=torch.float64, device="cuda")
        u = lambda *s: torch.rand(*s, generator=g, **f64) * 2 - 1
        self.sweeps = None
        self.ins = []
        if name in ("c1", "g4k"):
            n = 512 if name == "c1" else 4096
            self.n = n
            self.ins, self.out = [u(n, n), u(n, n)], [torch.empty(n, n, **f64)]
        elif name == "c2":
            self.ins, self.out = [u(units, 32, 32), u(units, 32, 32)], [torch.empty(units, 32, 32, **f64)]
        elif name == "c2b":
            self.ins, self.out = [u(units, 32, 32), u(1, 32, 32)], [torch.empty(units, 32, 32, **f64)]
        elif name == "c3":
            gmat = u(units, 16, 16)
            self.ins = [torch.baddbmm(16.0 * torch.eye(16, **f64).expand(units, 16, 16), gmat, gmat.transpose(1, 2))]
            del gmat
            self.out = [torch.empty(units, 16, 16, **f64)]
            self.info = torch.full((1,), 2 ** 62, dtype=torch.int64, device="cuda")
        elif name == "s3":
            gmat = u(units, 16, 16)
            low = torch.linalg.cholesky(torch.baddbmm(16.0 * torch.eye(16, **f64).expand(units, 16, 16), gmat, gmat.transpose(1, 2))).contiguous()  # row-major
            del gmat
            self.ins, self.out = [low, u(units, 16, 1)], [torch.empty(units, 16, 1, **f64)]
        elif name == "l4":
            q, r = torch.linalg.qr(u(units, 64, 32))
            self.ins, self.out = [q.contiguous(), r.contiguous(), u(units, 64, 1)], [torch.empty(units, 32, 1, **f64)]
        elif name == "i4":
            self.ins = [u(units, 64, 32), u(units, 64, 1)]
            self.out = [torch.empty(units, 64, 32, **f64), torch.empty(units, 64, 1, **f64)]
        elif name == "c4":
            self.ins = [u(units, 64, 32)]
            self.out = [torch.empty(units, 64, 32, **f64), torch.empty(units, 32, 32, **f64)]
        elif name == "v5":
            uu, ss, vv = torch.linalg.svd(u(units, 64, 64))
            self.ins = [uu.contiguous(), ss.contiguous(), vv.contiguous(), u(units, 64, 1)]
            del uu, ss, vv
            self.out = [torch.empty(units, 64, 1, **f64)]
            self.flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        elif name == "c5":
            self.ins = [u(units, 64, 64)]
            self.out = [torch.empty(units, 64, 64, **f64), torch.empty(units, 64, **f64), torch.empty(units, 64, 64, **f64)]
            self.sweeps = torch.zeros(4, dtype=torch.int32, device="cuda")
            # per-matrix sweep counts of every launch are summed here: work per matrix is proportional to its sweeps
            self.sweep_sum = torch.zeros(1, dtype=torch.int64, device="cuda")
            self.launches = 0
            self.pre_sum = torch.zeros(1, dtype=torch.int64, device="cuda")   # FP32 sweeps of the preconditioner
            if lib.nd4b_dev_svd_sweep_counter(dev, C.c_void_p(self.sweep_sum.data_ptr())) or \
                    lib.nd4b_dev_svd_pre_sweep_counter(dev, C.c_void_p(self.pre_sum.data_ptr())):
                raise RuntimeError(lib.nd4b_last_error().decode())
            ws = lib.nd4b_dev_svd_workspace(units, 64, 64)
            self.work = torch.empty(max(ws, 8) // 8, **f64)
            self.work_bytes = ws
        torch.cuda.synchronize()

    def mean_sweeps(self):
        """Mean Jacobi sweeps per matrix over all launches so far (None for the other workloads)."""
        if self.sweeps is None or not self.launches:
            return None
        return float(self.sweep_sum.item()) / (self.launches * self.units)

    def mean_pre_sweeps(self):
        if self.sweeps is None or not self.launches:
            return None
        return float(self.pre_sum.item()) / (self.launches * self.units)

    def plain_sweeps(self):
        """Mean sweeps of the plain FP64 iteration (no preconditioner: the launch gets no workspace) on the same input — the
        work a pure-FP64 one-sided Jacobi needs, for the 'equivalent' roofline figure."""
        L, p, t = self.lib, (lambda x: C.c_void_p(x.data_ptr())), self.torch
        tmp = t.zeros(1, dtype=t.int64, device="cuda")
        L.nd4b_dev_svd_sweep_counter(self.dev, p(tmp))
        i, o = self.ins, self.out
        rc = L.nd4b_dev_svd_jac1_f64(self.dev, C.c_void_p(t.cuda.current_stream().cuda_stream), p(i[0]), p(o[0]), p(o[1]), p(o[2]),
                                     self.units, 64, 64, None, None, 0)
        t.cuda.synchronize()
        L.nd4b_dev_svd_sweep_counter(self.dev, p(self.sweep_sum))
        if rc:
            raise RuntimeError(L.nd4b_last_error().decode())
        return float(tmp.item()) / self.units

    def close(self):
        if self.sweeps is not None:
            self.lib.nd4b_dev_svd_sweep_counter(self.dev, None)
            self.lib.nd4b_dev_svd_pre_sweep_counter(self.dev, None)

    def step(self, stream):
        L, p, d = self.lib, (lambda t: C.c_void_p(t.data_ptr())), self.dev
        s = C.c_void_p(stream)
        i, o = self.ins, self.out
        if self.name in ("c1", "g4k"):
            rc = L.nd4b_dev_matmul_f64(d, s, p(i[0]), 0, p(i[1]), 0, p(o[0]), 1, self.n, self.n, self.n)
        elif self.name == "c2":
            rc = L.nd4b_dev_matmul_f64(d, s, p(i[0]), 1024, p(i[1]), 1024, p(o[0]), self.units, 32, 32, 32)
        elif self.name == "c2b":
            rc = L.nd4b_dev_matmul_f64(d, s, p(i[0]), 1024, p(i[1]), 0, p(o[0]), self.units, 32, 32, 32)
        elif self.name == "c3":
            rc = L.nd4b_dev_cholesky_f64(d, s, p(i[0]), p(o[0]), self.units, 16, p(self.info))
        elif self.name == "s3":
            rc = L.nd4b_dev_tri_solve_f64(d, s, 2, p(i[0]), 256, p(i[1]), 16, p(o[0]), self.units, 16, 1)
        elif self.name == "l4":
            rc = L.nd4b_dev_qr_lstsq_f64(d, s, p(i[0]), p(i[1]), p(i[2]), p(o[0]), self.units, 64, 32, 32, 1)
        elif self.name == "i4":
            rc = L.nd4b_dev_qr_inplace_f64(d, s, p(i[0]), p(i[1]), p(o[0]), p(o[1]), self.units, 64, 32, 1)
        elif self.name == "c4":
            rc = L.nd4b_dev_qr_f64(d, s, p(i[0]), p(o[0]), p(o[1]), self.units, 64, 32, None, 0)
        elif self.name == "v5":
            rc = L.nd4b_dev_svd_lstsq_f64(d, s, p(i[0]), p(i[1]), p(i[2]), p(i[3]), p(o[0]), self.units, 64, 64, 64, 1, p(self.flag))
        else:
            rc = L.nd4b_dev_svd_jac1_f64(d, s, p(i[0]), p(o[0]), p(o[1]), p(o[2]), self.units, 64, 64, p(self.sweeps),
                                         p(self.work) if self.work_bytes else None, self.work_bytes)
            self.launches += 1
        if rc:
            raise RuntimeError("nd4b call failed: %s" % self.lib.nd4b_last_error().decode())

    def parity(self):
        """Checks a sample of what the last launch wrote against the CPU oracle."""
        t = self.torch
        t.cuda.synchronize()
        if self.name in ("c1", "g4k"):
            rows = t.as_tensor(sample_index(self.n, 4 if self.name == "g4k" else 171), device="cuda")
            return parity_of(self.name, [self.ins[0][rows].cpu().numpy(), self.ins[1].cpu().numpy()], [self.out[0][rows].cpu().numpy()])
        idx = t.as_tensor(sample_index(self.units, 8 if self.name == "c5" else 16), device="cuda")
        take = lambda x: (x if x.shape[0] == 1 else x[idx]).cpu().numpy()
        return parity_of(self.name, [take(x) for x in self.ins], [take(x) for x in self.out])


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


def host_case(name, nd, units, pinned=True):
    """The e2e leg: inputs as NDArrays in page-locked host memory (pinned=True; what a previous nd.la result is, or
    nd.pinned_array(x)) or as ordinary numpy arrays (pinned=False), and the PUBLIC operator nd4js_b200.la.<op> on them.
    Returns (call, h2d bytes, d2h bytes, sample) where call() runs one step and returns the result NDArray(s) and
    sample(result) gives the (inputs, outputs) of the parity check."""
    import numpy as np
    la = nd.la
    rng = np.random.default_rng(99)
    wrap = nd.pinned_array if pinned else (lambda x: np.ascontiguousarray(x))
    raw = lambda x: x.numpy() if hasattr(x, "numpy") else x
    u = lambda *s: rng.uniform(-1, 1, s)
    if name in ("c1", "c2", "c2b", "g4k"):
        sh = (512, 512) if name == "c1" else (4096, 4096) if name == "g4k" else (units, 32, 32)
        shb = (1, 32, 32) if name == "c2b" else sh
        ins = [wrap(u(*sh)), wrap(u(*shb))]
        call = lambda: la.matmul2(*ins)
    elif name == "c3":
        g = u(units, 16, 16)
        ins = [wrap(g @ g.transpose(0, 2, 1) + 16 * np.eye(16))]
        call = lambda: la.cholesky_decomp(*ins)
    elif name == "l4":
        qq, rr = np.linalg.qr(u(units, 64, 32))
        ins = [wrap(qq), wrap(rr), wrap(u(units, 64, 1))]
        call = lambda: la.qr_lstsq(*ins)
    elif name == "s3":
        g = u(units, 16, 16)
        ins = [wrap(np.linalg.cholesky(g @ g.transpose(0, 2, 1) + 16 * np.eye(16))), wrap(u(units, 16, 1))]
        call = lambda: la.cholesky_solve(*ins)
    elif name == "i4":
        ins = [wrap(u(units, 64, 32)), wrap(u(units, 64, 1))]
        call = lambda: la._qr_decomp_inplace(*ins)
    elif name == "c4":
        ins = [wrap(u(units, 64, 32))]
        call = lambda: la.qr_decomp(*ins)
    elif name == "v5":
        uu, ss, vv = np.linalg.svd(u(units, 64, 64))
        ins = [wrap(uu), wrap(ss), wrap(vv), wrap(u(units, 64, 1))]
        call = lambda: la.svd_lstsq(*ins)
    else:
        ins = [wrap(u(units, 64, 64))]
        call = lambda: la.svd_jac_1sided(*ins)
    first = call()
    outs = first if isinstance(first, tuple) else (first,)
    h2d = sum(raw(x).size for x in ins) * 8
    d2h = sum(o.data.size for o in outs) * 8

    def sample(result):
        res = result if isinstance(result, tuple) else (result,)
        if name in ("c1", "g4k"):
            n = raw(ins[0]).shape[0]
            rows = sample_index(n, 4 if name == "g4k" else 171)
            return [raw(ins[0])[rows], raw(ins[1])], [res[0].numpy()[rows]]
        idx = sample_index(units, 8 if name == "c5" else 16)
        take = lambda x: x if x.shape[0] == 1 else x[idx]
        return [take(raw(x)) for x in ins], [take(r.numpy()) for r in res]

    return call, h2d, d2h, sample


def time_host(call, steps, sync):
    t0 = time.perf_counter()
    res = None
    per = []
    for _ in range(steps):
        t1 = time.perf_counter()
        res = call()
        per.append(round(1e3 * (time.perf_counter() - t1), 2))
    sync()
    dt = time.perf_counter() - t0
    if os.environ.get("BENCH_DEBUG"):
        sys.stderr.write("time_host per-call ms: %s\n" % per)
    return dt, res


def fp64_peaks(lib, dev):
    """DFMA / DMMA pipe peaks measured on this GPU in this run (MEASURED_PEAKS.json has no fp64 entry)."""
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


def c5_sharded(torch, dist, lib, local, rank, world, max_over_ranks):
    """BASELINE config 5 as north_star states it: svd_jac_1sided of [16384,64,64] sharded over the ranks (contiguous
    split, no collective on the compute path), U / sv / V gathered with NCCL over NVLink.  Rank 0 checks every gathered
    matrix (residual, orthogonality, order) and compares a sample bit for bit with its own unsharded computation."""
    from nd4js_b200.partition import gather_shards, shard_range
    total = 16384
    g = torch.Generator(device="cuda").manual_seed(7)          # the same full batch on every rank: shards are views
    a = torch.rand(total, 64, 64, generator=g, dtype=torch.float64, device="cuda") * 2 - 1
    b0, b1 = shard_range(total, rank, world)
    mine = a[b0:b1].contiguous()
    f64 = dict(dtype=torch.float64, device="cuda")
    ws = lib.nd4b_dev_svd_workspace(total, 64, 64)
    work = torch.empty(max(ws, 8) // 8, **f64)
    p = lambda t: C.c_void_p(t.data_ptr())

    def svd(x):
        b = x.shape[0]
        u, sv, v = torch.empty(b, 64, 64, **f64), torch.empty(b, 64, **f64), torch.empty(b, 64, 64, **f64)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if lib.nd4b_dev_svd_jac1_f64(local, st, p(x), p(u), p(sv), p(v), b, 64, 64, None, p(work) if ws else None, ws):
            raise RuntimeError(lib.nd4b_last_error().decode())
        return u, sv, v

    [gather_shards(t, total) for t in svd(mine)]                 # warm-up of the kernel and of NCCL's channels
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    dist.barrier()
    torch.cuda.synchronize()
    e[0].record()
    u, sv, v = svd(mine)
    e[1].record()
    fu, fsv, fv = (gather_shards(t, total) for t in (u, sv, v))
    e[2].record()
    torch.cuda.synchronize()
    t_compute = max_over_ranks(e[0].elapsed_time(e[1]), device="cuda")
    t_gather = max_over_ranks(e[1].elapsed_time(e[2]), device="cuda")
    out = None
    if rank == 0:
        rec = (fu * fsv[:, None, :]) @ fv
        res = float(((rec - a).flatten(1).norm(dim=1) / a.flatten(1).norm(dim=1)).max())
        eye = torch.eye(64, **f64)
        orth = float(max((fu.transpose(1, 2) @ fu - eye).abs().max(), (fv @ fv.transpose(1, 2) - eye).abs().max()))
        sorted_ok = bool((fsv[:, :-1] >= fsv[:, 1:]).all() and (fsv >= 0).all())
        idx = torch.arange(0, total, total // 64, device="cuda")    # a sample across all shards, recomputed unsharded
        su, ssv, sv_ = svd(a[idx].contiguous())
        same = bool((su == fu[idx]).all() and (ssv == fsv[idx]).all() and (sv_ == fv[idx]).all())
        gathered = (fu.numel() + fsv.numel() + fv.numel()) * 8
        out = {"workload": "svd_jac_1sided [16384,64,64] sharded over %d GPUs, NCCL all-gather of U, sv, V" % world,
               "ms_compute": t_compute, "ms_gather_nccl": t_gather, "gather_gbs": gathered / t_gather / 1e6,
               "matrices_per_s_compute": total / t_compute * 1e3, "matrices_per_s_with_gather": total / (t_compute + t_gather) * 1e3,
               "max_rel_residual": res, "max_orth_error": orth, "sorted_nonneg": sorted_ok, "sharded_equals_unsharded_bits": same,
               "ok": bool(res <= 1e-12 and orth <= 1e-12 and sorted_ok and same)}
    del a, mine, fu, fsv, fv, u, sv, v, work
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import numpy as np
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
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its banner / debug lines to stdout by default; stdout carries ONE JSON line, so send them to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")   # host-side waits: no rank may spin a kernel on a GPU another rank is timing
    nd.init([local])
    lib = nd.load()
    desc, units, bpu, fpu = WORKLOADS[args.workload]
    hbm_peak, peak_src = measured_peaks()
    failures = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sync = torch.cuda.synchronize
    peaks = fp64_peaks(lib, local) if args.fp64_probes else {}
    fp64_peak = peaks.get("dmma_tflops", FP64_PEAK_TFLOPS)
    fp64_src = ("measured in this run (nd4b_probe_fp64: DMMA.8x8x4 %.2f, DFMA %.2f TFLOP/s; both share one pipe)"
                % (peaks.get("dmma_tflops", 0), peaks.get("dfma_tflops", 0))) if peaks else \
        "profiles/r01_fp64_peak.json (DMMA.8x8x4 37.15, DFMA 36.83 TFLOP/s on this pool's B200)"

    # ---------------- device-resident (value + roofline) ----------------
    case = DeviceCase(args.workload, torch, lib, local, units)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    secs = time_device(torch, case, args.steps, max(args.warmup, 3))
    barrier()
    secs = max_over_ranks(secs, device="cuda")
    # the same launch back to back for at least half a second: what the clocks settle to under a sustained load
    n_long = max(args.steps, int(args.sustained_seconds / max(secs / args.steps, 1e-6)) + 1)
    long_secs = max_over_ranks(time_device(torch, case, n_long, 0), device="cuda")
    sweeps = int(case.sweeps[0].item()) if case.sweeps is not None else None
    sweeps_mean = case.mean_sweeps()
    pre_mean = case.mean_pre_sweeps() if case.sweeps is not None else None
    value = world * units * args.steps / secs
    launch_s = secs / args.steps
    # Jacobi work is proportional to the sweeps each matrix needed; with the single-precision preconditioner only the FP64
    # sweeps and the four 64^3 FP64 products of the hand-over count as FP64 flop (the FP32 sweeps are reported beside them)
    flop_unit = fpu * (sweeps_mean if sweeps_mean else 1) + (4 * 2 * 64 ** 3 if pre_mean else 0)
    achieved_gbs = bpu * units / launch_s / 1e9
    parity = {"device_resident": case.parity()} if rank == 0 else {}
    plain_mean = case.plain_sweeps() if pre_mean else None   # after the parity check: it overwrites the outputs
    case.close()
    del case
    torch.cuda.empty_cache()

    if args.kernel_only:  # tuning aid: device-resident timing only, not a bench line
        if sampler:
            sampler.stop()
        if rank == 0:
            emit({"workload": args.workload, "ms_per_launch": 1e3 * launch_s, "ms_per_launch_sustained": 1e3 * long_secs / n_long,
                  "matrices_per_s": value, "gflops": value * flop_unit / 1e9, "hbm_gbs_algorithmic": achieved_gbs,
                  "sweeps_max": sweeps, "sweeps_mean": sweeps_mean, "fp32_pre_sweeps_mean": pre_mean,
                  "plain_fp64_sweeps_mean": plain_mean,
                  "frac_of_fp64_peak_equivalent": (value / world * fpu * plain_mean / 1e12 / fp64_peak) if plain_mean else None, "parity": parity})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- end to end through the operator (nd4js_b200.la.*), pinned and pageable ----------------
    call, h2d, d2h, sample = host_case(args.workload, nd, units, pinned=True)
    e2e_steps = max(1, min(args.steps, 10))
    last = None
    for _ in range(3):   # warm-up as the timed loop runs: the previous result is still referenced while the next call allocates
        last = call()    # its own, so the pinned-block cache ends up holding the two blocks the steady state alternates between
    last = None
    s0 = nd.stats()
    barrier()
    e2e_secs, last = time_host(call, e2e_steps, sync)
    barrier()
    e2e_secs = max_over_ranks(e2e_secs, device="cuda")
    s1 = nd.stats()
    e2e_value = world * units * e2e_steps / e2e_secs
    e2e_launches = s1["kernel_launches"] - s0["kernel_launches"]
    staged_pinned = s1["staged_bytes"] - s0["staged_bytes"]
    if rank == 0:
        ins_s, outs_s = sample(last)
        parity["e2e"] = parity_of(args.workload, ins_s, outs_s)
    del call, sample
    pcall, _, _, _ = host_case(args.workload, nd, units, pinned=False)
    last = pcall()
    last = pcall()
    last = None
    barrier()
    page_steps = max(1, min(args.steps, 3))
    page_secs, _ = time_host(pcall, page_steps, sync)
    barrier()
    page_secs = max_over_ranks(page_secs, device="cuda")
    del pcall
    clocks = sampler.stop() if sampler else None  # sampled over the timed regions so far (device-resident + e2e)

    line = None
    if rank == 0:
        traffic, traffic_src = NCU_TRAFFIC_BYTES.get(args.workload, (None, None))
        if args.workload in ("c1", "g4k", "c5"):
            # compute-bound configs: against the FP64 pipe — DMMA for the GEMMs, DFMA for the Jacobi SVD; both
            # instruction kinds share one pipe and one peak on B200
            tflops = value / world * flop_unit / 1e12
            roofline = {"bound": "tensor", "achieved": tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                        "frac": tflops / fp64_peak, "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": "FP64 pipe peak " + fp64_src,
                        "kernel": KERNEL[args.workload], "algorithmic_flop_per_unit": flop_unit, "units_per_launch": units}
            if plain_mean:
                # the preconditioner moves most sweeps to FP32, so the FP64 flop actually executed understate the work done:
                # `equivalent` charges the sweeps the plain FP64 iteration needs on the same input (measured in this run)
                eq = value / world * fpu * plain_mean / 1e12
                roofline["equivalent"] = {"plain_fp64_sweeps_mean": plain_mean, "tflops": eq, "frac": eq / fp64_peak,
                                          "note": "matrices/s x (flop per sweep x sweeps of the plain FP64 one-sided Jacobi on the same input)"}
        else:
            roofline = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved_gbs / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": peak_src, "kernel": KERNEL[args.workload], "algorithmic_bytes_per_unit": bpu, "units_per_launch": units}
        line = {
            "metric": "matrices/s", "value": value, "unit": "matrices/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * launch_s, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "units_per_gpu": units, "inputs": "U(-1,1) (SPD G.G^T+16I for cholesky), seeded",
                       "l2": "inputs+outputs per step (%.0f MiB) exceed the 126 MB L2; no explicit flush" % (bpu * units / 2 ** 20)
                       if bpu * units > 200e6 else "working set fits L2: kernel-only figure is L2-warm, see e2e"},
            "gflops": value * flop_unit / 1e9,
            "sustained": {"value": world * units * n_long / long_secs, "ms_per_step": 1e3 * long_secs / n_long, "launches": n_long,
                          "seconds": long_secs, "note": "the same launch back to back for >= %.2g s, CUDA events, max over ranks" % args.sustained_seconds},
            "e2e": {"value": e2e_value, "unit": "matrices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_secs / e2e_steps,
                    "timing": "host wall clock around the blocking operator call, barrier+sync both sides, max over ranks",
                    "api": "nd4js_b200.la.%s (NDArrays in page-locked host memory in and out; nd4b_*_f64 underneath)" % OPERATOR[args.workload],
                    "staged_bytes": staged_pinned, "rank_cpus_near_gpu": numa},
            "e2e_pageable": {"value": world * units * page_steps / page_secs, "unit": "matrices/s", "steps": page_steps,
                             "ms_per_step": 1e3 * page_secs / page_steps,
                             "api": "nd4js_b200.la.%s on ordinary numpy inputs (staged through the library's pinned ring); results pinned" % OPERATOR[args.workload]},
            "gpu_launches": args.steps + n_long + int(e2e_launches),
            "roofline": roofline,
            "clocks": clocks,
        }
        if sweeps:
            line["sweeps_max"], line["sweeps_mean"], line["fp32_pre_sweeps_mean"] = sweeps, sweeps_mean, pre_mean
        if peaks:
            line["fp64_peaks_measured"] = peaks

    # ---------------- the other BASELINE configs and 'next' rows, kernel-only, short ----------------
    if args.workload == "c2" and not args.no_others:
        others = {}
        for name in ("c2b", "c1", "g4k", "c3", "s3", "c4", "l4", "i4", "c5", "v5"):
            d2, u2, b2, f2 = WORKLOADS[name]
            c2 = DeviceCase(name, torch, lib, local, u2)
            barrier()
            n2 = 3 if name == "c5" else 10
            s2 = time_device(torch, c2, n2, 3)
            s2 = max_over_ranks(s2, device="cuda")
            sw = c2.mean_sweeps() or 1
            pre2 = c2.mean_pre_sweeps() if name == "c5" else None
            par2 = c2.parity() if rank == 0 else None   # before plain_sweeps(), which overwrites the outputs
            plain2 = c2.plain_sweeps() if pre2 else None
            gf = world * u2 * n2 / s2 * (f2 * sw + (4 * 2 * 64 ** 3 if pre2 else 0)) / 1e9
            gbs = b2 * u2 * n2 / s2 / 1e9
            others[name] = {"workload": d2, "matrices_per_s": world * u2 * n2 / s2, "ms_per_launch": 1e3 * s2 / n2,
                            "gflops": gf, "hbm_gbs_algorithmic": gbs,
                            "frac_of_hbm_peak": gbs / hbm_peak, "frac_of_fp64_peak": gf / world / 1e3 / fp64_peak,
                            "sweeps_mean": sw if name == "c5" else None, "fp32_pre_sweeps_mean": pre2,
                            "plain_fp64_sweeps_mean": plain2,
                            "frac_of_fp64_peak_equivalent": (u2 * n2 / s2 * f2 * plain2 / 1e12 / fp64_peak) if plain2 else None,
                            "sweeps_max": int(c2.sweeps[0].item()) if name == "c5" else None}
            if rank == 0:
                others[name]["parity"] = par2
                parity["others." + name] = par2
            c2.close()
            del c2
            torch.cuda.empty_cache()
        if rank == 0:
            line["others"] = others

    # ---------------- N > 1: the configuration north_star shards, and one call over all devices ----------------
    if world > 1 and not args.no_sharded:
        sh = c5_sharded(torch, dist, lib, local, rank, world, max_over_ranks)
        if rank == 0:
            line["c5_sharded"] = sh
            if not sh["ok"]:
                failures.append("c5_sharded")
        # strong scaling of ONE operator call: rank 0 alone re-creates its context over all N devices and runs the e2e leg
        # through it; the other ranks wait on the host (gloo) so that nothing of theirs runs on the GPUs being timed
        dist.barrier(group=cpu_group)
        if rank == 0:
            # the same call through the one-device context first, with the other ranks idle: the strong-scaling reference
            call1, _, _, _ = host_case(args.workload, nd, units, pinned=True)
            last = call1()
            last = call1()
            last = None
            t_one, last = time_host(call1, e2e_steps, lambda: None)
            del call1
            last = None
            nd._lib.check(lib.nd4b_shutdown())
            nd.init(list(range(world)))
            call, h2d_s, d2h_s, sample = host_case(args.workload, nd, units, pinned=True)
            last = call()
            last = call()
            last = None
            t_sh, last = time_host(call, e2e_steps, lambda: None)
            ins_s, outs_s = sample(last)
            par = parity_of(args.workload, ins_s, outs_s)
            parity["e2e_sharded"] = par
            line["e2e_sharded"] = {"value": units * e2e_steps / t_sh, "unit": "matrices/s", "n_devices": nd.stats()["n_devices"],
                                   "ms_per_step": 1e3 * t_sh / e2e_steps, "steps": e2e_steps,
                                   "one_device_idle_box": {"value": units * e2e_steps / t_one, "ms_per_step": 1e3 * t_one / e2e_steps},
                                   "speedup_vs_one_device_call": t_one / t_sh,
                                   "note": "one nd4js_b200.la.%s call of %d units from rank 0, its context over all %d devices "
                                           "(contiguous shards, no collective), against the same call through a one-device context "
                                           "with the other ranks idle: strong scaling of a single call" % (OPERATOR[args.workload], units, world)}
            del call, sample, last
        dist.barrier(group=cpu_group)

    # ---------------- CPU baseline: rank 0, N=1 only ----------------
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import nd4ref
        nd4ref.build()
        n_ref = REF_SAMPLE[args.workload]
        fn, data = _ref_case(args.workload, n_ref, np.random.default_rng(3), nd4ref)
        reps, t0 = 0, time.perf_counter()
        while True:
            fn(*data)
            reps += 1
            if time.perf_counter() - t0 > 10.0 or reps >= 50:
                break
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n_ref * reps / dt, "unit": "matrices/s", "cores": 1, "kind": "port",
                                "sample": "%d units x %d passes (%.1f s) with oracle/libnd4ref.so, a C restatement of the JS loops "
                                          "(no JS engine in this image); host has %d cores, the reference is single-threaded"
                                          % (n_ref, reps, dt, os.cpu_count())}
    if rank == 0:
        line["parity"] = parity
        failures += [k for k, v in parity.items() if not v["ok"]]
        if failures:
            # a fast kernel whose results differ from the reference's is not done: no value is reported
            sys.stderr.write("bench.py: PARITY FAILED for %s: %s\n" % (failures, json.dumps(parity)))
            emit({"error": "parity check failed", "failed": failures, "parity": parity, "c5_sharded": line.get("c5_sharded")})
            if world > 1:
                dist.destroy_process_group()
            raise SystemExit(3)
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
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the c5_sharded and e2e_sharded legs")
    ap.add_argument("--sustained-seconds", type=float, default=0.5,
                    help="length of the back-to-back timing next to the K-step figure (shorten it for an ncu launch list)")
    ap.add_argument("--kernel-only", action="store_true", help="tuning aid: print the device-resident timing only")
    ap.add_argument("--no-fp64-probes", dest="fp64_probes", action="store_false",
                    help="do not measure the DFMA/DMMA pipe peaks in this run (use the recorded 37.15 TFLOP/s)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
