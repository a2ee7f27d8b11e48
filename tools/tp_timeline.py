"""Per-launch phase timeline of the peer-memory TP step on rank 0 (needs -DFP4_STREAM_TIMELINE build).
torchrun --nproc-per-node 2 tools/tp_timeline.py"""
import ctypes
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
from torch_bnb_fp4_b200._lib import lib  # noqa: E402
from torch_bnb_fp4_b200.graph import GraphedCallable  # noqa: E402
from torch_bnb_fp4_b200.parallel import PeerExchange  # noqa: E402

cfg = dict(bench.MISTRAL); cfg["layers"] = 3
layers, _ = bench.build_stack(cfg, dev, rank, world)
h0 = torch.randn(1, cfg["hidden"], device=dev).bfloat16()
ex = PeerExchange(cfg["hidden"], torch.bfloat16, dev)
step = bench.make_step_peer(layers, ex)
KW, NL = 16, 4 * cfg["layers"]
stride = 148 * KW * 8
buf = torch.zeros(stride * (NL * 6 + 8), dtype=torch.int64, device=dev)
with torch.no_grad():
    step(h0)
torch.cuda.synchronize()
g = GraphedCallable(step, [h0], warmup=2)
torch.cuda.synchronize(); dist.barrier()
lib.fp4_b200_debug_stream_timeline.argtypes = [ctypes.c_void_p]
# stamps are indexed by launch order since the call: re-capture so the graph's launches get slots 0..NL-1
lib.fp4_b200_debug_stream_timeline(buf.data_ptr())
g2 = GraphedCallable(step, [h0], warmup=1)   # warm-up launch uses slots 0..NL-1, capture NL..2NL-1
for _ in range(3):
    g2.graph.replay()
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    raw = buf.cpu().numpy().reshape(-1, 148 * KW, 8)
    first = NL  # the captured launches
    t00 = raw[first:first + NL][raw[first:first + NL] > 0].min()
    names = ["start", "slot0 issued", "dep-wait done", "x staged", "first data", "loop done", "(bar)", "end"]
    kinds = ["qkv", "o (produce)", "gate/up (consume)", "down (produce)"]
    for li in range(NL):
        r = raw[first + li]
        print(f"launch {li} {kinds[li % 4]}: active warps {(r[:, 0] > 0).sum()}")
        for j, nm in enumerate(names):
            col = r[:, j][r[:, j] > 0]
            if col.size and j != 6:
                print(f"   {nm:14s} min {col.min() - t00:8d}  max {col.max() - t00:8d} ns")
torch.cuda.synchronize(); sys.stdout.flush()
dist.barrier()
os._exit(0)
