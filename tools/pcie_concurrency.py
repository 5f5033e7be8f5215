"""PCIe / host-memory concurrency of the box: k = 1, 2, 4, ... GPUs each stream the C2 traffic pattern (1 GiB H2D and
0.5 GiB D2H per step, page-locked host memory, the two directions on separate streams) AT THE SAME TIME, one worker process
per GPU as in `bench.py --gpus N`.  Prints the per-GPU and aggregate rates: the cause of the e2e scaling of the bench
(each GPU alone is PCIe-bound; together they share the host's memory system and PCIe root complexes).

  python tools/pcie_concurrency.py            # sweeps k over the visible GPUs
"""
import json
import os
import subprocess
import sys
import time


def worker(dev, k, reps, start_at):
    import torch
    torch.cuda.set_device(dev)
    try:   # the CPUs next to this GPU, as bench.py does
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        cpus = set()
    n = 1 << 27
    hin = torch.empty(n, dtype=torch.float64).pin_memory()
    hin.fill_(1.0)
    hout = torch.empty(n // 2, dtype=torch.float64).pin_memory()
    din = torch.empty(n, dtype=torch.float64, device="cuda")
    dout = torch.ones(n // 2, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def step():
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)

    step()
    torch.cuda.synchronize()
    while time.time() < start_at:   # all workers start together
        pass
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"dev": dev, "k": k, "ms_per_step": 1e3 * dt, "h2d_gbs": 2 ** 30 / dt / 1e9, "d2h_gbs": 2 ** 29 / dt / 1e9,
                      "cpus": len(cpus)}))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        worker(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]))
        return
    import torch
    n_gpu = torch.cuda.device_count()
    print("GPUs visible: %d, host CPUs: %d" % (n_gpu, os.cpu_count()))
    k = 1
    while k <= n_gpu:
        start_at = time.time() + 25.0
        procs = [subprocess.Popen([sys.executable, __file__, "--worker", str(d), str(k), "8", repr(start_at)], stdout=subprocess.PIPE, text=True)
                 for d in range(k)]
        rows = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
        agg_in, agg_out = sum(r["h2d_gbs"] for r in rows), sum(r["d2h_gbs"] for r in rows)
        slow = max(r["ms_per_step"] for r in rows)
        print("k=%d concurrent GPUs: per-GPU step %.1f .. %.1f ms; aggregate H2D %.1f GB/s + D2H %.1f GB/s; C2-equivalent %.2f M matrices/s (ideal x%d: %.2f)"
              % (k, min(r["ms_per_step"] for r in rows), slow, agg_in, agg_out, k * 65536 / slow / 1e3, k, k * 65536 / rows[0]["ms_per_step"] / 1e3 if k == 1 else 0))
        k *= 2


if __name__ == "__main__":
    main()
