"""Multi-GPU sharding of the path (one process per GPU, `torch.distributed`): SURVEY.md §8e.

* the z independent OU processes (S/models.py:118, batch dimension of the MVN) are split into contiguous z-slices:
  rank r runs assembly / Cholesky / gradients only for its slice;
* the decoder / encoder GRUs are data-parallel over contiguous blocks of leaves;
* parameters are replicated; per step there is ONE all-reduce (flattened gradients + the scalar loss) and, for the
  variational guide, one all-gather of `[z_loc | z_scale]` over the leaf blocks (its backward is a reduce-scatter).
  These are latency-bound messages (<= a few hundred KB), so they go through NCCL rather than a fused kernel.
* sites whose log-density is replicated on every rank (hyper-priors, q(z)) are kept on rank 0 only (scale 0 elsewhere),
  so that the SUM over ranks of the local losses is exactly the single-process loss.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def split_even(total: int, parts: int) -> List[Tuple[int, int]]:
    """Contiguous near-equal ranges: 30 over 4 -> 8,8,7,7; 30 over 8 -> 4,4,4,4,4,4,3,3."""
    base, extra = divmod(total, parts)
    out, lo = [], 0
    for r in range(parts):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


class _AllGatherRows(torch.autograd.Function):
    """Concatenate per-rank row blocks (uneven allowed); backward sums the incoming gradient over ranks and returns
    this rank's block (reduce-scatter)."""

    @staticmethod
    def forward(ctx, x, shard):
        ctx.shard = shard
        pad = shard.max_leaf_block
        buf = x.new_zeros((pad,) + tuple(x.shape[1:]))
        buf[: x.shape[0]] = x
        outs = [torch.empty_like(buf) for _ in range(shard.world)]
        dist.all_gather(outs, buf.contiguous(), group=shard.group)
        return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, shard.leaf_ranges)], dim=0)

    @staticmethod
    def backward(ctx, g):
        shard = ctx.shard
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=shard.group)
        lo, hi = shard.leaf_ranges[shard.rank]
        return g[lo:hi], None


class ShardContext:
    def __init__(self, rank: int, world: int, z_dim: int, n_leaves: int, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.z_ranges = split_even(z_dim, world)
        self.leaf_ranges = split_even(n_leaves, world)
        self.z_lo, self.z_hi = self.z_ranges[rank]
        self.leaf_lo, self.leaf_hi = self.leaf_ranges[rank]
        self.max_leaf_block = max(hi - lo for lo, hi in self.leaf_ranges)
        # replicated log-densities are counted once: on rank 0
        self.replicated_scale = 1.0 if rank == 0 else 0.0

    def gather_leaf_rows(self, x_local: torch.Tensor) -> torch.Tensor:
        return _AllGatherRows.apply(x_local, self)

    def gather_z_rows(self, x_local: torch.Tensor) -> torch.Tensor:
        """Concatenate the ranks' z-slices (dimension 0, uneven allowed) of a tensor without gradient: the conditional
        posterior means `[z_r, ni]` / draws `[z_r, ni, S]` of the z-sharded K6.  One all-gather."""
        pad = max(hi - lo for lo, hi in self.z_ranges)
        buf = x_local.new_zeros((pad,) + tuple(x_local.shape[1:]))
        buf[: x_local.shape[0]] = x_local
        outs = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(outs, buf.contiguous(), group=self.group)
        return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, self.z_ranges)], dim=0)

    def broadcast_from_rank0(self, x: torch.Tensor) -> torch.Tensor:
        """Every rank continues with rank 0's tensor (random numbers that all ranks must agree on)."""
        dist.broadcast(x, src=0 if self.group is None else dist.get_global_rank(self.group, 0), group=self.group)
        return x

    def reduce_step(self, params: List[torch.Tensor], loss: torch.Tensor, flag=None):
        """One all-reduce(SUM) over [every gradient | loss | pivot-failure count]; returns (global loss, global failure
        count) — detached device scalars, so that every rank sees a failed factorisation of any rank's z-slice."""
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
        fl = torch.zeros(1, dtype=loss.dtype, device=loss.device) if flag is None else flag.detach().reshape(1).to(loss.dtype)
        flat = torch.cat([g.reshape(-1) for g in grads] + [loss.detach().reshape(1), fl])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        off = 0
        for p, g in zip(params, grads):
            k = g.numel()
            p.grad = flat[off: off + k].view_as(g).clone() if p.grad is None else p.grad.copy_(flat[off: off + k].view_as(g))
            off += k
        return flat[off], flat[off + 1]
