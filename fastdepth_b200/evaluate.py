"""Image-sharded evaluation: every rank runs the forward on its own images, per-image metrics
are summed on device, and ONE all-reduce(SUM) of 11 doubles finishes the job.

Reproduces the reference's validate() bookkeeping (main.py:63-127): ``Result.evaluate`` per
image at batch size 1 (main.py:40-41, 80-82) accumulated by ``AverageMeter`` (metrics.py:71-95),
i.e. the MEAN OF PER-IMAGE metrics -- not metrics of pooled pixels (SURVEY.md section 8e trap).
The reference itself is single-process; the sharding is new and has exactly one collective.
"""
import torch
import torch.distributed as dist

from .plan import METRIC_NAMES, N_METRICS, metrics_accumulate


def shard_range(n_total, rank, world_size):
    """Contiguous image range [lo, hi) of ``rank``; remainders go to the lowest ranks."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def new_sums(device):
    return torch.zeros(N_METRICS, dtype=torch.float64, device=device)


def reduce_sums(sums, group=None):
    """The path's single collective: all-reduce(SUM) of [10 metric sums, count] in fp64
    (NCCL on GPUs over NVLink/NVSwitch -- 88 bytes, latency only; gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def finalize(sums):
    """AverageMeter.average() (reference metrics.py:84-95): sums / count."""
    s = sums.detach().cpu().tolist()
    count = s[N_METRICS - 1]
    out = {k: (s[i] / count if count else float('nan')) for i, k in enumerate(METRIC_NAMES)}
    out['count'] = count
    return out


@torch.no_grad()
def evaluate(model, batches, device, group=None, return_sums=False):
    """``batches`` yields (input [b,3,H,W], target [b,1,H,W]) for THIS rank's shard (host or
    device tensors).  Returns the averaged metrics over all ranks' images (with ``return_sums`` also the reduced
    11-double sum vector, for bookkeeping checks)."""
    sums = new_sums(device)
    model.eval()
    dtype = next(model.parameters()).dtype
    for inp, tgt in batches:
        inp = inp.to(device=device, dtype=dtype, non_blocking=True)
        tgt = tgt.to(device=device, dtype=torch.float32, non_blocking=True)
        pred = model(inp)
        metrics_accumulate(pred, tgt, sums)
    reduce_sums(sums, group)
    return (finalize(sums), sums) if return_sums else finalize(sums)
