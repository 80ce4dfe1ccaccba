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
def evaluate(model, batches, device, group=None, return_sums=False, lanes=1):
    """``batches`` yields (input [b,3,H,W], target [b,1,H,W]) for THIS rank's shard (host or
    device tensors).  Returns the averaged metrics over all ranks' images (with ``return_sums`` also the reduced
    11-double sum vector, for bookkeeping checks).

    ``lanes`` > 1 keeps that many batches in flight (``engine.ForwardLanes``: plan copies on their own streams; each lane
    accumulates into its own sum vector on its own stream, the vectors are added at the end).  The result does not depend on
    it: every per-image metric enters the fp64 sums as an fp32-rounded value, so the additions are exact in any order."""
    model.eval()
    dtype = next(model.parameters()).dtype
    if lanes > 1:
        from .engine import ForwardLanes
        fl = ForwardLanes(model, lanes=lanes)
        lane_sums = [new_sums(device) for _ in range(lanes)]
        streams = fl.streams_for(torch.device(device))
        for inp, tgt in batches:
            inp = inp.to(device=device, dtype=dtype, non_blocking=True)
            tgt = tgt.to(device=device, dtype=torch.float32, non_blocking=True)
            lane = fl.next
            pred, _ = fl.forward(inp)                       # the lane's stream first waits for the current stream (inp, tgt ready)
            tgt.record_stream(streams[lane])
            with torch.cuda.stream(streams[lane]):
                metrics_accumulate(pred, tgt, lane_sums[lane])
        fl.synchronize()
        sums = lane_sums[0]
        for other in lane_sums[1:]:
            sums = sums + other
    else:
        sums = new_sums(device)
        for inp, tgt in batches:
            inp = inp.to(device=device, dtype=dtype, non_blocking=True)
            tgt = tgt.to(device=device, dtype=torch.float32, non_blocking=True)
            pred = model(inp)
            metrics_accumulate(pred, tgt, sums)
    reduce_sums(sums, group)
    return (finalize(sums), sums) if return_sums else finalize(sums)
