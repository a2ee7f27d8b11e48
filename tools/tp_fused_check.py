"""2-rank check of the peer-memory tensor-parallel exchange against the NCCL path and the unsharded layers.
torchrun --nproc-per-node 2 tools/tp_fused_check.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
from torch_bnb_fp4_b200.graph import GraphedCallable  # noqa: E402
from torch_bnb_fp4_b200.parallel import PeerExchange  # noqa: E402

cfg = dict(hidden=2048, inter=4096, kv=1024, layers=3)
tp_layers, _ = bench.build_stack(cfg, dev, rank, world)
full_layers, _ = bench.build_stack(cfg, dev, 0, 1)
h0 = torch.randn(1, cfg["hidden"], device=dev, generator=torch.Generator(device=dev).manual_seed(5)).bfloat16()
def say(msg):
    torch.cuda.synchronize()
    print(f"[rank {rank}] {msg}", flush=True)


with torch.no_grad():
    say("stacks built")
    ref_full = bench.make_step(full_layers, 1)(h0).float()
    say("full done")
    ref_nccl = bench.make_step(tp_layers, world)(h0).float()
    say("nccl done")
    ex = PeerExchange(cfg["hidden"], torch.bfloat16, dev)
    say("exchange ready")
    step_p = bench.make_step_peer(tp_layers, ex)
    outs = []
    for i in range(3):
        outs.append(step_p(h0).float().clone())
        say(f"peer step {i} done, state {ex.state.tolist()}")
    ex.check()
    emu = bench.make_step_peer(tp_layers, ex, emulate=True)(h0).float()
scale = ref_full.abs().max().item()
print(f"[rank {rank}] peer == all_gather + fp32 rank-order sum, bit for bit: {torch.equal(emu, outs[0])}", flush=True)
print(f"[rank {rank}] nccl vs full {((ref_nccl - ref_full).abs().max() / scale).item():.2e}  "
      f"peer vs full {((outs[0] - ref_full).abs().max() / scale).item():.2e}  "
      f"peer run-to-run {((outs[2] - outs[0]).abs().max() / scale).item():.2e}", flush=True)
allout = [torch.empty_like(outs[0]) for _ in range(world)]
dist.all_gather(allout, outs[0])
print(f"[rank {rank}] ranks agree bit for bit: {all(torch.equal(allout[0], o) for o in allout)}", flush=True)
# varying batch sizes with the ranks deliberately out of step (the consumer waits for late peers; a peer that
# never delivers traps after 20 s instead of being summed as garbage)
import time  # noqa: E402
worst = 0.0
with torch.no_grad():
    step_n = bench.make_step(tp_layers, world)
    step_e = bench.make_step_peer(tp_layers, ex, emulate=True)
    exact = True
    for it, b in enumerate((1, 3, 2, 8, 1, 5)):
        xb = torch.randn(b, cfg["hidden"], device=dev, generator=torch.Generator(device=dev).manual_seed(40 + it)).bfloat16()
        if it % 2 == rank % 2:
            torch.cuda.synchronize()
            time.sleep(0.05)
        got = step_p(xb).float()
        want = step_n(xb).float()
        worst = max(worst, ((got - want).abs().max() / want.abs().max()).item())
        exact = exact and torch.equal(got, step_e(xb).float())
    ex.check()
print(f"[rank {rank}] skewed ranks, batch 1..8: bit-identical to the all_gather emulation: {exact}", flush=True)
print(f"[rank {rank}] skewed ranks, batch 1..8: peer vs nccl worst {worst:.2e}", flush=True)
g = GraphedCallable(step_p, [h0], warmup=3)
for _ in range(5):
    og = g(h0).float()
ex.check()
print(f"[rank {rank}] graph replay vs eager {((og - outs[0]).abs().max() / scale).item():.2e}", flush=True)
# past 2^16 epochs (the 16-bit tags of round 1 came round again there): a batch-8 step, 14 000 batch-1 replays
# (5 epochs each), then batch 8 again with the ranks out of step - words left by the first batch-8 step must not validate
with torch.no_grad():
    x8 = torch.randn(8, cfg["hidden"], device=dev, generator=torch.Generator(device=dev).manual_seed(77)).bfloat16()
    want8 = step_n(x8).float()
    step_p(x8)
    for _ in range(14000):
        g(h0)
    if rank == 0:
        torch.cuda.synchronize()
        time.sleep(0.05)
    got8 = step_p(x8).float()
    ex.check()
print(f"[rank {rank}] after {int(ex.state[1].item())} epochs: batch 8 peer vs nccl "
      f"{((got8 - want8).abs().max() / want8.abs().max()).item():.2e}", flush=True)
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
