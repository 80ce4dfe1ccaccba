"""world_size-2 gloo test of the sharded-evaluation bookkeeping: shards +
one all-reduce of 11 doubles must reproduce the single-process per-image averages exactly."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastdepth_b200 import evaluate
from fastdepth_b200.plan import METRIC_NAMES, N_METRICS
from oracle import fastdepth_oracle as orc


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _data(n=7):
    rng = np.random.Generator(np.random.PCG64(11))
    out = (rng.random((n, 1, 16, 24), dtype=np.float32) * 4 + 0.3).astype(np.float32)
    tgt = (out * (1 + 0.15 * rng.standard_normal(out.shape).astype(np.float32))).clip(1e-3, None).astype(np.float32)
    return out, tgt


def _sums(out, tgt):
    s = torch.zeros(N_METRICS, dtype=torch.float64)
    for o, t in zip(out, tgt):
        r = orc.evaluate_one(o, t)
        for i, k in enumerate(METRIC_NAMES):
            s[i] += r[k]
        s[N_METRICS - 1] += 1
    return s


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    out, tgt = _data()
    lo, hi = evaluate.shard_range(len(out), rank, world)
    s = evaluate.reduce_sums(_sums(out[lo:hi], tgt[lo:hi]))
    if rank == 0:
        q.put(s.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    for n in (1, 7, 64, 65, 512):
        for world in (1, 2, 3, 8):
            spans = [evaluate.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_two_rank_reduce_matches_single_process():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = torch.tensor(q.get(timeout=120), dtype=torch.float64)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    out, tgt = _data()
    want = _sums(out, tgt)
    assert torch.allclose(got, want, rtol=1e-12, atol=0)
    avg = evaluate.finalize(got)
    ref, n = orc.average_per_image(out, tgt)
    assert avg['count'] == n == 7
    for k in METRIC_NAMES:
        assert abs(avg[k] - ref[k]) <= 1e-9 * max(1.0, abs(ref[k]))
