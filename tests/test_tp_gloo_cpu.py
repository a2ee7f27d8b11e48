"""world_size-2 gloo test of the tensor-parallel sharding rules: shards are cut
straight out of the bitsandbytes buffers, so a column shard must dequantise to exactly the rows of the
unsharded weight, and the row-parallel partial sums must all-reduce to the unsharded result.  Compute
is done by the CPU oracle here; the same helpers feed the CUDA kernels on GPUs (tests/test_gpu_tp.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, K, bs):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from helpers import synth_quant
    from torch_bnb_fp4_b200.parallel import shard_column, shard_row

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    packed, absmax, _ = synth_quant(N * K, bs, seed=7)
    x = np.random.default_rng(8).standard_normal((3, K)).astype(np.float32)
    code = oracle.bnb_code()
    full_w = oracle.dequant_tree(packed, absmax, N * K, bs, oracle.BF16).reshape(N, K)
    full_y = oracle.linear_f64(x, packed, absmax, code, None, N, K, bs)

    # column parallel: bit-identical shard, all_gather reproduces the full output
    p, a, n = shard_column(torch.from_numpy(packed), torch.from_numpy(absmax), N, K, rank, world, bs)
    w_shard = oracle.dequant_tree(p.numpy().ravel(), a.numpy(), n * K, bs, oracle.BF16).reshape(n, K)
    assert np.array_equal(w_shard, full_w[rank * n:(rank + 1) * n])
    y_local = torch.from_numpy(oracle.linear_f64(x, p.numpy().ravel(), a.numpy(), code, None, n, K, bs))
    parts = [torch.empty_like(y_local) for _ in range(world)]
    dist.all_gather(parts, y_local)
    assert np.array_equal(torch.cat(parts, dim=-1).numpy(), full_y)

    # row parallel: shard = column slice of every row; partial sums all-reduce to the full output
    p, a, k = shard_row(torch.from_numpy(packed), torch.from_numpy(absmax), N, K, rank, world, bs)
    w_shard = oracle.dequant_tree(p.numpy().ravel(), a.numpy(), N * k, bs, oracle.BF16).reshape(N, k)
    assert np.array_equal(w_shard, full_w[:, rank * k:(rank + 1) * k])
    y_part = torch.from_numpy(oracle.linear_f64(x[:, rank * k:(rank + 1) * k], p.numpy().ravel(), a.numpy(),
                                                code, None, N, k, bs))
    dist.all_reduce(y_part)
    assert np.allclose(y_part.numpy(), full_y, rtol=1e-12, atol=1e-12)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N,K,bs", [(64, 256, 64), (96, 512, 128)])
def test_column_and_row_sharding_world2(N, K, bs):
    mp.spawn(_worker, args=(2, _free_port(), N, K, bs), nprocs=2, join=True)


def test_shard_validation():
    from torch_bnb_fp4_b200.parallel import shard_column, shard_row

    packed = torch.zeros(64 * 192 // 2, dtype=torch.uint8)
    absmax = torch.zeros(64 * 192 // 64)
    with pytest.raises(ValueError):
        shard_row(packed, absmax, 64, 192, 0, 2, 64)     # 96 columns per rank: not a block multiple
    with pytest.raises(ValueError):
        shard_column(packed, absmax, 64, 192, 0, 3, 64)  # 64 rows not divisible by 3
    p, a, n = shard_column(packed, absmax, 64, 192, 1, 2, 64)
    assert n == 32 and p.numel() == 32 * 96 and a.numel() == 32 * 3
