"""CUDA-graph replay of a stack of FP4 layers.

Every kernel of this package launches on the current stream and never synchronises, so a decode step
(a fixed sequence of ``TorchFP4Linear`` calls) can be captured once and replayed: at batch 1 a layer
takes a few microseconds on B200, well below the cost of launching it from Python.  The reference
cannot do this (its kernels are hard-wired to the legacy default stream,
reference csrc/gemv_fp4_optimized.cu:266).
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedCallable:
    """Capture ``fn(*inputs)`` once; ``__call__`` copies new inputs in and replays.

    ``inputs`` may live on the host (pinned memory recommended): they are copied into the captured
    static input buffers with non-blocking copies on the current stream.  The returned tensors are
    the captured static outputs (valid until the next call)."""

    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 3):
        self.static_inputs = [t.detach().clone() for t in example_inputs]
        self.graph = torch.cuda.CUDAGraph()
        stream = torch.cuda.Stream(device=self.static_inputs[0].device)
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream), torch.no_grad():
            for _ in range(max(1, warmup)):  # also creates per-stream workspaces before capture
                fn(*self.static_inputs)
            stream.synchronize()
            with torch.cuda.graph(self.graph, stream=stream):
                out = fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(stream)
        self.static_outputs = out

    def __call__(self, *inputs: torch.Tensor):
        for dst, src in zip(self.static_inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_outputs
