"""Multi-process host logic on CPU (gloo, world size 2): the only collective on
the fwd+loss path is the validation-epoch mean of `log_dict(sync_dist=True)`
(reference base_model.py:84); batches are sharded by shape with no exchange."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from multi_part_assembly_b200.configs import get_cfg
    from multi_part_assembly_b200.datasets import make_batch
    from multi_part_assembly_b200.models import build_model
    model = build_model(get_cfg('pn_transformer'))
    assert model.local_rank == rank and model.global_rank == rank
    # validation_epoch_end: batch-size weighted mean per rank, then all-reduce mean
    outputs = [{'loss': torch.tensor(1.0 + rank), 'part_acc': torch.tensor(0.5 * rank), 'batch_size': 4},
               {'loss': torch.tensor(3.0 + rank), 'part_acc': torch.tensor(0.5 * rank), 'batch_size': 12}]
    model.validation_epoch_end(outputs)
    # shard a global batch of shapes across ranks: disjoint, covering, no communication
    global_batch = make_batch(8, P=20, N=16, num_valid=3, seed=0)
    shard = {k: v[rank::world] for k, v in global_batch.items()}
    ids = shard['data_id'].clone()
    gathered = [torch.zeros_like(ids) for _ in range(world)]
    dist.all_gather(gathered, ids)
    if rank == 0:
        out.put({k: float(v) for k, v in model._logged.items()})
        out.put(sorted(torch.cat(gathered).tolist()))
    dist.destroy_process_group()


def test_sync_dist_mean_and_sharding():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    logged = q.get(timeout=120)
    ids = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # rank r: (1+r)*4 + (3+r)*12 over 16 = 2.5 + r ; mean over ranks = 3.0
    assert abs(logged['val/loss'] - 3.0) < 1e-6
    assert abs(logged['val/part_acc'] - 0.25) < 1e-6
    assert ids == list(range(8))


def _grad_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from multi_part_assembly_b200.runtime import allreduce_gradients
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))  # same init on both ranks
    # a parameter that only rank 0 uses in this step: rank 1 has no gradient for it and must
    # contribute zeros, so that both ranks reduce buffers of the same length
    unused = torch.nn.Parameter(torch.zeros(4))
    x = torch.full((2, 5), float(rank + 1))
    loss = net(x).sum()
    if rank == 0:
        loss = loss + 3.0 * unused.sum()
    loss.backward()
    local = [p.grad.clone() for p in net.parameters()]
    allreduce_gradients(list(net.parameters()) + [unused])
    gathered = [[torch.zeros_like(g) for _ in range(world)] for g in local]
    for g, bufs in zip(local, gathered):
        dist.all_gather(bufs, g)
    ok = all(torch.allclose(p.grad, torch.stack(bufs).mean(0), atol=1e-6)
             for p, bufs in zip(net.parameters(), gathered)) and \
        unused.grad is not None and torch.allclose(unused.grad, torch.full((4, ), 1.5))
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_gradient_allreduce_averages_over_ranks():
    """Data-parallel training: one flat all-reduce averages every gradient (the
    reference's implicit DDP reduction, scripts/train.py:85,141)."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
