"""N > 1 host logic on CPU: two gloo ranks exercise the gradient-bucket all-reduce, parameter broadcast, running-stat
averaging and the max-over-ranks timing reduction used by bench.py."""
import os
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from phc_b200.learning import dist as D
    assert D.is_multi() and D.world_size() == world and D.rank_seed(5) == 5 + rank
    g = torch.full((1000,), float(rank + 1))
    scale = D.allreduce_grad_bucket(g)
    # the pipelined form the agent uses: begin() issues the collective (off the compute stream on CUDA), the caller does
    # weight-independent work of the next minibatch, end() joins; on CPU tensors it degenerates to the blocking call
    red = D.GradReducer("cpu")
    g2 = torch.full((1000,), float(2 * rank + 1))
    s2 = red.begin(g2)
    overlap_work = torch.arange(10).sum()          # stands for _prepare_minibatch(i + 1)
    red.end()
    assert s2 == 0.5 and float((g2 * s2)[0]) == 2.0 and int(overlap_work) == 45
    p = torch.full((10,), float(rank))
    D.broadcast_params(p)
    rms = types.SimpleNamespace(running_mean=torch.full((4,), float(rank), dtype=torch.float64),
                                running_var=torch.full((4,), 1.0 + rank, dtype=torch.float64),
                                count=torch.tensor(10.0 * (rank + 1), dtype=torch.float64))
    D.sync_running_stats([rms, None])
    t = D.max_over_ranks(1.0 + rank, "cpu")
    q.put((rank, float((g * scale)[0]), float(p[0]), float(rms.running_mean[0]), float(rms.running_var[0]), float(rms.count), t))
    dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gmean, p0, m, v, c, t in res:
        assert gmean == 1.5            # mean of the per-rank gradients
        assert p0 == 0.0               # parameters follow rank 0
        assert m == 0.5 and v == 1.5 and c == 15.0
        assert t == 2.0                # slowest rank defines the step time


def test_single_process_is_noop():
    from phc_b200.learning import dist as D
    g = torch.ones(8)
    assert D.allreduce_grad_bucket(g) == 1.0 and D.world_size() == 1 and not D.is_multi()
    red = D.GradReducer("cpu")
    assert red.begin(g) == 1.0 and red.end() is None and float(g.sum()) == 8.0
    assert D.max_over_ranks(3.0, "cpu") == 3.0
