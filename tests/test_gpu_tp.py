"""Tensor-parallel exchange through peer memory (torch_bnb_fp4_b200.parallel.PeerExchange): needs >= 2 GPUs of one
box (skipped otherwise).  Runs tools/tp_fused_check.py under torchrun: the peer-memory step must agree with the
NCCL step and the unsharded layers, be bit-identical across ranks and replay-stable under CUDA graphs."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_exchange_two_ranks():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "tp_fused_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=400, cwd=ROOT).stdout
    m = re.findall(r"nccl vs full ([0-9.e+-]+)\s+peer vs full ([0-9.e+-]+)\s+peer run-to-run ([0-9.e+-]+)", out)
    assert len(m) == 2, out
    for nccl, peer, rr in m:
        assert float(peer) <= 2e-2 and float(peer) <= 2.0 * float(nccl) + 1e-3 and float(rr) == 0.0
    assert out.count("ranks agree bit for bit: True") == 2
    assert out.count("peer == all_gather + fp32 rank-order sum, bit for bit: True") == 2, out
    assert out.count("bit-identical to the all_gather emulation: True") == 2, out
    skew = re.findall(r"skewed ranks, batch 1\.\.8: peer vs nccl worst ([0-9.e+-]+)", out)
    assert len(skew) == 2 and all(float(v) <= 3e-2 for v in skew), out
    assert len(re.findall(r"graph replay vs eager 0\.00e\+00", out)) == 2
    wrap = re.findall(r"after (\d+) epochs: batch 8 peer vs nccl ([0-9.e+-]+)", out)
    assert len(wrap) == 2 and all(int(e) > 65536 and float(v) <= 3e-2 for e, v in wrap), out


@pytest.mark.gpu
def test_one_process_two_devices():
    """the opt-in to large dynamic shared memory is per device; a process that drives two GPUs (accelerate's
    device_map) must be able to launch every kernel on both."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    import torch_bnb_fp4
    from torch_bnb_fp4_b200 import bnb_compat
    torch.manual_seed(0)
    w = torch.randn(1024, 2048) * 0.03
    for idx in (0, 1, 0):
        dev = torch.device("cuda", idx)
        m = torch_bnb_fp4.TorchFP4Linear(bnb_compat.make_quantized_linear(w.to(dev)))
        for rows in (1, 6, 64, 1500):
            x = torch.randn(rows, 2048, device=dev, dtype=torch.bfloat16)
            y = m(x)
            ref = torch.nn.functional.linear(x.float(), m.quant_data.dequantize().float())
            assert y.device == dev
            assert ((y.float() - ref).abs().max() / ref.abs().max()).item() <= 8e-3, (idx, rows)
