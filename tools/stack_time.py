"""Development: device time per decode token of the Mistral-7B (or other) linear stack, CUDA-graph replays,
ungrouped (7 launches per layer) and grouped (4).  Prints one line per mode; used for same-box A/B of kernel variants
(FP4_B200_LIB=...).  Not the judged bench (that is bench.py)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from torch_bnb_fp4_b200.graph import GraphedCallable  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mistral7b")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--tag", default=os.environ.get("FP4_B200_LIB", "default").split("libfp4_b200.")[-1])
    a = ap.parse_args()
    cfg = dict(bench.MISTRAL if a.workload == "mistral7b" else bench.LLAMA70B)
    if a.layers:
        cfg["layers"] = a.layers
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    layers, nbytes = bench.build_stack(cfg, dev)
    h0 = torch.randn(1, cfg["hidden"], device=dev).bfloat16()
    env = " ".join(f"{k[13:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("FP4_B200_GEMV_"))
    import torch_bnb_fp4
    blocks = [bench.Block(dict(m)) for m in layers]
    for b in blocks:
        torch_bnb_fp4.group_projections(b)
    for mode, stepfn in (("ungrouped", bench.make_step(layers, 1)), ("grouped", bench.make_step_blocks(blocks))):
        r = GraphedCallable(stepfn, [h0], warmup=3)
        for _ in range(5):
            r.graph.replay()
        best = 1e9
        for _ in range(3):
            best = min(best, bench.time_events(r.graph.replay, a.steps) / a.steps)
        print(f"[{a.tag} {env}] {a.workload} {mode:9s}: {best * 1e3:7.3f} ms/token  {nbytes / best / 1e9:7.1f} GB/s  "
              f"{1 / best:6.1f} tok/s", flush=True)
        del r


if __name__ == "__main__":
    main()
