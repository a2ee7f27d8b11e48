"""Tensor-parallel sharding of bitsandbytes-FP4 linears.

The reference has no distributed code at all; the north star adds a column-parallel split of
``out_features`` (q/k/v/gate/up) and a row-parallel split of ``in_features`` (o/down) over the GPUs of
one box, one process per GPU, NCCL over NVLink.

Sharding acts directly on the bitsandbytes buffers - no re-quantisation, so a shard dequantises to
exactly the rows/columns of the unsharded weight:

* column-parallel: rank r owns rows [r*N/tp, (r+1)*N/tp): a CONTIGUOUS slice of the flat packed
  bytes and of the flat absmax (needs (N/tp * K) % blocksize == 0);
* row-parallel: rank r owns columns [r*K/tp, (r+1)*K/tp) of every row: a strided slice of the
  [N, K/2] packed view and of the [N, K/blocksize] absmax view, made contiguous once at load
  (needs K % blocksize == 0 and (K/tp) % blocksize == 0);
* a nested (double-quantised) absmax is materialised to fp32 before slicing: its 256-block grouping
  does not line up with shard boundaries.

The collectives are plain ``torch.distributed`` calls (NCCL on GPUs; the sharding logic itself is
backend independent and is tested with gloo on CPU).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn


def shard_column(packed: torch.Tensor, absmax: torch.Tensor, N: int, K: int, rank: int, tp: int,
                 blocksize: int = 64) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Rows [rank*N/tp, (rank+1)*N/tp).  Returns (packed_shard [n*K/2, 1], absmax_shard, n_rows)."""
    if N % tp:
        raise ValueError(f"out_features {N} not divisible by tp {tp}")
    n = N // tp
    if (n * K) % blocksize or (n * K) % 2:
        raise ValueError("shard boundary falls inside a quantisation block")
    e0 = rank * n * K
    p = packed.reshape(-1)[e0 // 2:(e0 + n * K) // 2].contiguous().view(-1, 1)
    a = absmax.reshape(-1)[e0 // blocksize:(e0 + n * K) // blocksize].contiguous()
    return p, a, n


def shard_row(packed: torch.Tensor, absmax: torch.Tensor, N: int, K: int, rank: int, tp: int,
              blocksize: int = 64) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Columns [rank*K/tp, (rank+1)*K/tp) of every row.  Returns (packed_shard, absmax_shard, k_cols)."""
    if K % tp:
        raise ValueError(f"in_features {K} not divisible by tp {tp}")
    k = K // tp
    if K % blocksize or k % blocksize:
        raise ValueError(f"row-parallel shards need K and K/tp to be multiples of blocksize {blocksize}")
    p = packed.reshape(N, K // 2)[:, rank * k // 2:(rank + 1) * k // 2].contiguous().view(-1, 1)
    a = absmax.reshape(N, K // blocksize)[:, rank * k // blocksize:(rank + 1) * k // blocksize].contiguous().view(-1)
    return p, a, k


def _group_info(group) -> Tuple[int, int]:
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


class _ShardedFP4Base(nn.Module):
    def __init__(self, layer, mode: str, group=None, rank: Optional[int] = None, tp: Optional[int] = None):
        super().__init__()
        from . import TorchFP4Linear, _ext
        from .bnb_compat import LinearFP4, Params4bit, QuantState

        if not isinstance(layer, TorchFP4Linear):
            raise TypeError("expected a TorchFP4Linear")
        g_rank, g_tp = _group_info(group)
        self.rank = g_rank if rank is None else rank
        self.tp = g_tp if tp is None else tp
        self.group = group
        qd = layer.quant_data
        N, K, bs = qd.M, qd.N, qd.blocksize
        absmax = qd.absmax
        if absmax is None:  # nested: materialise before slicing
            absmax = _ext.absmax_denest(qd.nested, (N * K + bs - 1) // bs, qd.A.device)
        if mode == "column":
            p, a, n = shard_column(qd.A, absmax, N, K, self.rank, self.tp, bs)
            shape = (n, K)
            bias = None if qd.bias is None else qd.bias.detach()[self.rank * n:(self.rank + 1) * n].clone()
        else:
            p, a, k = shard_row(qd.A, absmax, N, K, self.rank, self.tp, bs)
            shape = (N, k)
            # the bias is added exactly once: by rank 0, before the reduction
            bias = qd.bias.detach().clone() if (qd.bias is not None and self.rank == 0) else None
        lin = LinearFP4(shape[1], shape[0], bias=bias is not None)
        st = QuantState(absmax=a, shape=shape, code=qd.code, blocksize=bs, quant_type="fp4",
                        dtype=getattr(qd.quant_state, "dtype", torch.float16))
        lin.weight = Params4bit(p, requires_grad=False, quant_state=st, blocksize=bs,
                                compress_statistics=False, quant_type="fp4")
        if bias is not None:
            lin.bias = nn.Parameter(bias, requires_grad=False)
        self.local = TorchFP4Linear(lin, use_codebook_dequant=layer.use_codebook_dequant,
                                    name=layer.name + f".tp{self.rank}")
        self.in_features, self.out_features = K, N


class ColumnParallelFP4Linear(_ShardedFP4Base):
    """y_local = x @ W[rows of this rank]^T.  ``gather_output=True`` all-gathers along features;
    leave it False when the consumer is a row-parallel layer (Megatron style, no collective)."""

    def __init__(self, layer, gather_output: bool = False, group=None, rank=None, tp=None):
        super().__init__(layer, "column", group, rank, tp)
        self.gather_output = gather_output

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.local(x)
        if not self.gather_output or self.tp == 1:
            return y
        parts = [torch.empty_like(y) for _ in range(self.tp)]
        dist.all_gather(parts, y.contiguous(), group=self.group)
        return torch.cat(parts, dim=-1)


class RowParallelFP4Linear(_ShardedFP4Base):
    """y = sum_ranks x[..., cols of this rank] @ W[:, cols]^T.  ``input_is_parallel=True`` means x is
    already the local slice (the output of a column-parallel layer)."""

    def __init__(self, layer, input_is_parallel: bool = True, group=None, rank=None, tp=None):
        super().__init__(layer, "row", group, rank, tp)
        self.input_is_parallel = input_is_parallel

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.input_is_parallel:
            k = self.in_features // self.tp
            x = x[..., self.rank * k:(self.rank + 1) * k]
        y = self.local(x)
        if self.tp > 1:
            dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)
        return y


class PeerExchange:
    """Tensor-parallel exchange through peer (symmetric) memory, shared by all row-parallel layers of a model
    (include/fp4_b200.h: fp4_b200_tp_t).  A row-parallel layer PUSHES its partial output, as self-validating
    64-bit words {two 16-bit values, tag32 = epoch}, into every rank's exchange buffer over NVLink (no fences, no
    flags, no collective launch); the next column-parallel layer sums the ranks' partials (fp32, rank order) out of
    its LOCAL buffer while it stages x.  Requires torch symmetric memory (NVLink peers of one box), an initialised
    process group and a 16-bit activation dtype."""

    def __init__(self, max_features: int, dtype: torch.dtype, device: torch.device, group=None, max_batch: int = 8):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib

        if dtype not in (torch.float16, torch.bfloat16):
            raise ValueError("PeerExchange carries 16-bit activations")
        group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("PeerExchange supports up to 8 ranks")
        self.dtype, self.device = dtype, device
        self.slot_words = ((max_features * max_batch // 2 + 63) // 64) * 64  # one rank's region of one slot
        self.slot_bytes = self.slot_words * 8                                # a word: {two 16-bit values, tag32}
        self.buf = symm_mem.empty(2 * self.world * self.slot_words, dtype=torch.int64, device=device)
        self.buf.zero_()
        self._hbuf = symm_mem.rendezvous(self.buf, group=group)
        self.state = torch.zeros(4, dtype=torch.int32, device=device)  # epochs[0], epochs[1], err, pad
        torch.cuda.synchronize(device)
        self._hbuf.barrier()
        self._lib = _lib

    def struct(self, consume: bool, produce: bool):
        t = self._lib.TpExchange()
        t.slot_bytes = self.slot_bytes
        t.epochs = self.state.data_ptr()
        t.err = self.state.data_ptr() + 8
        if consume:
            t.in_world = self.world
            t.in_base = self.buf.data_ptr()
        if produce:
            t.out_world, t.out_rank = self.world, self.rank
            for q in range(self.world):
                t.out_peer_base[q] = self._hbuf.buffer_ptrs[q]
        return t

    def check(self):
        """Host-side check (synchronises): raises if a kernel gave up waiting for a peer."""
        if int(self.state[2].item()) != 0:
            raise RuntimeError("tensor-parallel exchange: a wait for a peer's partial sums timed out")


class PeerPartial:
    """Handle for a row-parallel layer's partial output that lives in the PeerExchange (not a tensor)."""

    def __init__(self, exchange: "PeerExchange", batch_shape, features: int):
        self.exchange, self.batch_shape, self.features = exchange, tuple(batch_shape), features


def fused_tp_group_forward(layers, x, exchange: "PeerExchange", produce: bool = False):
    """Run TorchFP4Linear `layers` (same in_features) in one grouped launch with the peer-memory exchange:
    x may be a tensor or a PeerPartial (summed over ranks while staging x); with produce=True (one layer only)
    the output is published as a PeerPartial instead of being returned as a tensor."""
    from . import _ext

    qds = [m.quant_data for m in layers]
    consume = isinstance(x, PeerPartial)
    tp = exchange.struct(consume, produce)
    dt = exchange.dtype
    for q in qds:
        if q.o_type != dt:
            q.set_compute_type(torch.empty(0, dtype=dt))
    lead = x.batch_shape if consume else tuple(x.shape[:-1])
    outs = None
    if produce:
        if len(layers) != 1:
            raise ValueError("a producer is a single row-parallel layer")
        outs = [exchange.buf]
    res = _ext.gemv_fp4_grouped(None if consume else x.contiguous(), [q.A for q in qds], [q.absmax for q in qds], 64,
                                qds[0].qtype, [q._Bshape for q in qds], [q._bias_t for q in qds], tp=tp, outs=outs,
                                batch_shape=lead)
    if res is None:
        raise RuntimeError("shapes outside the grouped streaming kernel: use the NCCL path")
    if produce:
        return PeerPartial(exchange, lead, qds[0].M)
    return tuple(res)
