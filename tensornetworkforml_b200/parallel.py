"""Host-side pieces of the sample-sharded (data-parallel) sweep -- SURVEY.md section 8e.

Every Ns-proportional object (phi, environments, f, q) is per-sample, so rank r simply owns the contiguous sample
range ``shard_bounds(Ns, r, world)``.  Samples only mix in the gradient sum over b (NC:710), the two metrics
(NC:697-702) and the calibration maximum (NC:169): one all-reduce(sum) of ``[dB | n_correct | sum|y-f| | count]``
per bond update, one all-reduce(max) at construction.  Everything after the all-reduce (regularisation, clipping,
SVD split) is replicated on every rank and deterministic, so no broadcast is needed.

These helpers take plain torch tensors (CUDA + NCCL in production, CPU + gloo in the unit tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

N_EXTRA = 4   # n_correct, sum |y - f|, number of samples, sum |f| (debug history)


def shard_bounds(Ns: int, rank: int, world: int):
    """Contiguous, balanced sample range [lo, hi) of ``rank``; the union over ranks is [0, Ns) without overlap."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of size %d" % (rank, world))
    return rank * Ns // world, (rank + 1) * Ns // world


def reduce_gradient_and_metrics(buf: torch.Tensor, n_grad: int, local_count: int, group=None, world: int = 1,
                                count_written: bool = False):
    """``buf[:n_grad]`` = local dB, ``buf[n_grad]`` = local n_correct, ``buf[n_grad+1]`` = local sum|y-f|.
    Writes the local sample count into ``buf[n_grad+2]`` and sums the first ``n_grad + N_EXTRA`` entries over the
    group in place (a single collective per bond update)."""
    if not count_written:        # tnml_act_lossder writes [.., count, 0] itself: no extra launches on the critical path
        buf[n_grad + 2:n_grad + N_EXTRA].zero_()
        buf[n_grad + 2:n_grad + 3].fill_(float(local_count))
    if world > 1:
        dist.all_reduce(buf[:n_grad + N_EXTRA], op=dist.ReduceOp.SUM, group=group)
    return buf


def metrics_from_sums(n_correct, abs_err_sum, count, n_labels):
    """accuracy (NC:700) and MAE (NC:702) of the WHOLE batch from the reduced sums."""
    return n_correct / count, abs_err_sum / (count * n_labels)


def global_abs_max(local_max: float, device, group=None, world: int = 1) -> float:
    """max |f| over all shards: the calibration rescale of NC:169-170 must see the whole batch."""
    if world <= 1:
        return float(local_max)
    t = torch.tensor([float(local_max)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
