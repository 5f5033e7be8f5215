"""CPU: host-side mirror of the reference interface (NDArray, argument checks, chain ordering) and
the batch partitioner, including a world_size-2 gloo run of the sharded path's bookkeeping."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ndarray_contract():
    from nd4js_b200 import NDArray, asarray, from_numpy
    a = from_numpy(np.arange(6.0).reshape(2, 3))
    assert a.shape.dtype == np.int32 and list(a.shape) == [2, 3] and a.ndim == 2 and a.dtype == "float64"
    with pytest.raises(ValueError):
        a.shape[0] = 5  # frozen, nd_array.js:142
    with pytest.raises(ValueError, match="Shape must be Int32Array"):
        NDArray(np.array([2, 3]), np.zeros(6))
    with pytest.raises(ValueError, match="Invalid shape"):
        NDArray(np.array([0, 3], np.int32), np.zeros(0))
    with pytest.raises(ValueError, match="does not match"):
        NDArray(np.array([2, 3], np.int32), np.zeros(5))
    assert asarray(a) is a
    assert asarray([[1, 2], [3, 4]]).dtype == "int32"
    assert asarray([[1.5]]).dtype == "float64"
    t = a.T
    assert list(t.shape) == [3, 2] and (t.numpy() == a.numpy().T).all()


def test_argument_errors_carry_the_reference_texts():
    from nd4js_b200 import la
    with pytest.raises(ValueError, match="A must be at least 2D."):
        la.matmul2([1.0, 2.0], [[1.0], [2.0]])
    with pytest.raises(ValueError, match="B must be at least 2D."):
        la.matmul2([[1.0, 2.0]], [1.0, 2.0])
    with pytest.raises(ValueError, match="do not match"):
        la.matmul2(np.ones((2, 3)), np.ones((4, 2)))
    with pytest.raises(ValueError, match="broadcast-compatible"):
        la.matmul2(np.ones((2, 4, 3)), np.ones((3, 3, 2)))
    with pytest.raises(ValueError, match="quadratic"):
        la.cholesky_decomp(np.ones((2, 3)))
    with pytest.raises(ValueError, match="qr_decomp\\(A\\): A.ndim must be at least 2."):
        la.qr_decomp([1.0, 2.0])
    with pytest.raises(TypeError):
        la.cholesky_decomp(np.ones((2, 2), np.float32))


def test_matmul_chain_uses_the_flop_optimal_order():
    """matmul.js:185-235: for [10,2]x[2,10]x[10,2] the right product must be formed first; the plan is the postfix form
    handed to nd4b_matmul_plan_f64 (i pushes operand i, -1 multiplies the two topmost items)."""
    from nd4js_b200 import la

    def run(plan, mats):  # the plan's semantics, in NumPy
        st = []
        for op in plan:
            if op >= 0:
                st.append(mats[op])
            else:
                b_ = st.pop()
                a_ = st.pop()
                st.append(a_ @ b_)
        assert len(st) == 1
        return st[0]

    a, b, c = np.ones((10, 2)), np.ones((2, 10)), np.ones((10, 2))
    plan, shape = la._chain_plan([a.shape, b.shape, c.shape])
    assert plan == [0, 1, 2, -1, -1] and shape == [10, 2]
    plan, shape = la._chain_plan([(2, 10), (10, 2), (2, 10)])
    assert plan == [0, 1, -1, 2, -1] and shape == [2, 10]
    # broadcast-aware flop counts (matmul.js:159-180): a batched left factor makes the right-to-left order cheaper
    plan, shape = la._chain_plan([(7, 4, 3), (3, 50), (50, 2)])
    assert plan == [0, 1, 2, -1, -1] and shape == [7, 4, 2]
    g = np.load(os.path.join(ROOT, "tests", "golden", "known_answers.npz"))
    mats = [g["chain_a"], g["chain_b"], g["chain_c"]]
    plan, _ = la._chain_plan([m.shape for m in mats])
    assert (run(plan, mats) == g["chain_abc"]).all()
    rng = np.random.default_rng(0)
    for _ in range(20):  # any plan is a valid parenthesisation of the whole chain with the minimal flop count
        n = int(rng.integers(3, 7))
        dims = rng.integers(1, 9, n + 1)
        mats = [rng.integers(-3, 4, (int(dims[i]), int(dims[i + 1]))).astype(np.float64) for i in range(n)]
        plan, shape = la._chain_plan([m.shape for m in mats])
        assert sorted(p for p in plan if p >= 0) == list(range(n)) and plan.count(-1) == n - 1
        want = mats[0]
        for m in mats[1:]:
            want = want @ m
        assert (run(plan, mats) == want).all() and list(want.shape) == shape
    with pytest.raises(ValueError, match="Shape mismatch."):
        la._chain_plan([(2, 3), (4, 5), (5, 6)])
    assert la.matmul(a) is not None


def test_shard_ranges_cover_the_batch():
    sys.path.insert(0, ROOT)
    from nd4js_b200.partition import shard_range
    for total in (1, 7, 8, 65536, 262144 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_sharding_reassembles_the_reference_result(ref, tmp_path):
    """world_size 2 on CPU (gloo): each rank takes its shard of a cholesky batch, the shard results
    (computed by the oracle standing in for the device, which is absent here) are gathered and must
    equal the unsharded result; the timing reduction (max over ranks) is exercised too."""
    script = tmp_path / "rank.py"
    script.write_text(
        "import os, sys, numpy as np, torch, torch.distributed as dist\n"
        "sys.path.insert(0, %r)\n"
        "from nd4js_b200.partition import shard_range, gather_shards, max_over_ranks\n"
        "from oracle import nd4ref\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "g = np.random.default_rng(0).uniform(-1, 1, (37, 8, 8)); s = g @ g.transpose(0, 2, 1) + 8 * np.eye(8)\n"
        "b0, b1 = shard_range(37, r, w)\n"
        "mine = nd4ref.cholesky_decomp(s[b0:b1])\n"
        "full = gather_shards(torch.from_numpy(mine), 37)\n"
        "t = max_over_ranks(float(r + 1))\n"
        "e0, e1 = shard_range(36, r, w)\n"
        "even = gather_shards(torch.from_numpy(np.ascontiguousarray(s[e0:e1])), 36)   # equal shards: one all_gather_into_tensor\n"
        "if r == 0:\n"
        "    assert (full.numpy() == nd4ref.cholesky_decomp(s)).all(); assert t == float(w)\n"
        "    assert (even.numpy() == s[:36]).all(); print('OK')\n"
        "dist.destroy_process_group()\n" % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK" in out.stdout


def test_svd_rank_and_lstsq_have_no_cpu_path():
    """svd_rank / svd_lstsq run behind the C ABI on the device (src/la/svd.js:31-226): without a usable GPU they fail loudly —
    argument errors with the reference's texts are still raised first, on the host."""
    import torch
    from nd4js_b200 import Nd4bError, la
    with pytest.raises(ValueError, match="U and y don't match"):
        la.svd_lstsq(np.eye(3), np.ones(3), np.eye(3), np.ones((4, 1)))
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the device path is covered by the -m gpu tests")
    with pytest.raises(Nd4bError, match="no usable CUDA device|no CPU fallback"):
        la.svd_rank(np.array([[3.0, 1.0, 1e-9, 0.0]]))
    with pytest.raises(Nd4bError, match="no usable CUDA device|no CPU fallback"):
        la.svd_lstsq(np.eye(3), np.ones(3), np.eye(3), np.ones((3, 1)))
