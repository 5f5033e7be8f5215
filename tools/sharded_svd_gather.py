"""BASELINE config 5 as north_star states it: svd_jac_1sided of [16384,64,64] sharded over the GPUs of one box (one process
per GPU, contiguous split of the batch index, no collective on the compute path), results gathered with NCCL.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29571 \
      tools/sharded_svd_gather.py [batch]

Rank 0 checks the gathered [U, sv, V] against the defining properties on every matrix and, bit for bit, against its own
unsharded computation of a sample, and prints one JSON line with the compute and gather times (CUDA events, max over ranks)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import nd4js_b200 as nd  # noqa: E402
from nd4js_b200.partition import shard_range, gather_shards, max_over_ranks  # noqa: E402


def svd_device(lib, dev, a):
    b = a.shape[0]
    f64 = dict(dtype=torch.float64, device=a.device)
    u, sv, v = torch.empty(b, 64, 64, **f64), torch.empty(b, 64, **f64), torch.empty(b, 64, 64, **f64)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.nd4b_dev_svd_jac1_f64(dev, st, p(a), p(u), p(sv), p(v), b, 64, 64, None, None, C.c_size_t(0))
    if rc:
        raise RuntimeError(lib.nd4b_last_error().decode())
    return u, sv, v


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nd.init([local])
    lib = nd.load()
    g = torch.Generator(device="cuda").manual_seed(7)                  # the same full batch on every rank: shards are views
    a = torch.rand(total, 64, 64, generator=g, dtype=torch.float64, device="cuda") * 2 - 1
    b0, b1 = shard_range(total, rank, world)
    mine = a[b0:b1].contiguous()
    svd_device(lib, local, mine[: min(8, b1 - b0)])                       # warm-up
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e[0].record()
    u, sv, v = svd_device(lib, local, mine)
    e[1].record()
    fu, fsv, fv = (gather_shards(t, total) for t in (u, sv, v))
    e[2].record()
    torch.cuda.synchronize()
    t_compute = max_over_ranks(e[0].elapsed_time(e[1]), device="cuda")
    t_gather = max_over_ranks(e[1].elapsed_time(e[2]), device="cuda")
    if rank == 0:
        rec = (fu * fsv[:, None, :]) @ fv
        res = float(((rec - a).flatten(1).norm(dim=1) / a.flatten(1).norm(dim=1)).max())
        eye = torch.eye(64, dtype=torch.float64, device="cuda")
        orth = float(max((fu.transpose(1, 2) @ fu - eye).abs().max(), (fv @ fv.transpose(1, 2) - eye).abs().max()))
        sorted_ok = bool((fsv[:, :-1] >= fsv[:, 1:]).all() and (fsv >= 0).all())
        idx = torch.arange(0, total, max(1, total // 64), device="cuda")     # a sample across all shards, recomputed unsharded
        su, ssv, sv_ = svd_device(lib, local, a[idx].contiguous())
        same = bool((su == fu[idx]).all() and (ssv == fsv[idx]).all() and (sv_ == fv[idx]).all())
        gathered = (fu.numel() + fsv.numel() + fv.numel()) * 8
        print(json.dumps({"workload": "svd_jac_1sided [%d,64,64] over %d GPU(s)" % (total, world), "ms_compute": t_compute,
                          "ms_gather_nccl": t_gather, "gather_gbs": gathered / t_gather / 1e6 if world > 1 else None,
                          "matrices_per_s_compute": total / t_compute * 1e3, "matrices_per_s_with_gather": total / (t_compute + t_gather) * 1e3,
                          "max_rel_residual": res, "max_orth_error": orth, "sorted_nonneg": sorted_ok,
                          "sharded_equals_unsharded_bits": same}))
        assert res <= 1e-12 and orth <= 1e-12 and sorted_ok and same
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
