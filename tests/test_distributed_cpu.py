"""Host logic of the N > 1 path (joint-vae_b200/distributed.py) on CPU with gloo, world_size 2: parameter broadcast at
attach, the single averaged all-reduce hooked into the optimizer, sample sharding for scoring and the final gather.
The arithmetic kernels need a GPU; what is checked here is the plumbing every rank runs around them."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    try:
        sys.path.insert(0, ROOT)
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
        import __graft_entry__ as g
        pkg = g.load_package()
        torch.manual_seed(100 + rank)                    # different initial weights on every rank
        net = pkg.ClassificationVariationalNetwork((1, 8, 8), 4, type='cvae', encoder=[16], decoder=[16], classifier=[],
                                                   latent_dim=4, latent_sampling=2, gamma=0,
                                                   prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar'},
                                                   sigma={'value': 0.5})
        pkg.distributed.attach(net, bf16_bucket=True)
        # 1. every rank holds rank 0's parameters after attach
        flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], t) for t in gathered), 'parameters differ after attach'
        # 2. the hooked all-reduce averages the flat gradient bucket
        gbuf = torch.full((flat.numel(),), float(rank + 1))
        red = net.optimizer.allreduce(gbuf)
        assert torch.allclose(red, torch.full_like(red, sum(range(1, world + 1)) / world))
        # 3. scoring shards cover [0, n) once and the gather restores the sample order on rank 0
        n = 11
        lo, hi = pkg.distributed.shard_range(n)
        ranges = [None] * world
        dist.all_gather_object(ranges, (lo, hi))
        assert ranges[0][0] == 0 and ranges[-1][1] == n and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        local = {'elbo': torch.arange(lo, hi, dtype=torch.float32), 'iws': -torch.arange(lo, hi, dtype=torch.float32)}
        full = pkg.distributed.gather_scores(local, n, dst=0)
        if rank == 0:
            assert torch.equal(full['elbo'], torch.arange(n, dtype=torch.float32))
            assert torch.equal(full['iws'], -torch.arange(n, dtype=torch.float32))
        else:
            assert full is None
        dist.barrier()
        dist.destroy_process_group()
        out[rank] = 'ok'
    except Exception as e:      # noqa: BLE001 - reported to the parent
        out[rank] = f'{type(e).__name__}: {e}'


def test_data_parallel_plumbing_world_size_2():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: 'ok', 1: 'ok'}, dict(out)
