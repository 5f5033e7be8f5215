"""Runs one device-resident workload a few times (for ncu / quick timing): python tools/run_case.py c5 [units] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import nd4js_b200 as nd  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1]
units = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[name][1]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.cuda.set_device(0)
nd.init([0])
case = bench.DeviceCase(name, torch, nd.load(), 0, units)
secs = bench.time_device(torch, case, reps, 2)
print("%s units=%d: %.4f ms per launch, %.3e units/s" % (name, units, 1e3 * secs / reps, units * reps / secs))
