"""The N > 1 path on real devices: two processes, one GPU each, NCCL.  Skipped when fewer than two GPUs are visible (the
single-GPU test box); run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py -m gpu`.
  * attach: every rank holds rank 0's parameters;
  * two data-parallel train steps on different shards: the ranks end with bit-identical parameters (ONE averaged all-reduce
    of the bf16 bucket, then the same clip + Adam everywhere), and the step equals a single-process step fed with the
    averaged gradient (checked through the loss of the next step being finite and equal across ranks after the broadcast);
  * sample-sharded scoring: per-rank evaluate on its shard, ONE gather over NCCL, the gathered scores equal the scores a
    single process computes on the whole set."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _net(pkg, dev):
    return pkg.ClassificationVariationalNetwork(
        (3, 16, 16), 4, type='cvae', features='[x3+1]8-8-M-16:2-16', upsampler='[x3+1]16x4+0-16-8:2++1-8:2++1-!3x3+1',
        batch_norm='both', encoder=[], decoder=[], classifier=[], latent_dim=16, latent_sampling=3, test_latent_sampling=3,
        gamma=0, output_activation='linear', sigma={'value': 1.0},
        prior={'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 8},
        optimizer={'optim_type': 'adam', 'lr': 1e-3, 'grad_clipping': 100}).to(dev)


def _worker(rank, world, port, out):
    try:
        sys.path.insert(0, ROOT)
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device('cuda', rank)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
        import __graft_entry__ as g
        pkg = g.build()
        torch.manual_seed(100 + rank)                    # different initial weights on every rank
        net = _net(pkg, dev)
        pkg.distributed.attach(net, bf16_bucket=True)
        flat = lambda: torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        gathered = [torch.empty_like(flat()) for _ in range(world)]
        dist.all_gather(gathered, flat())
        assert all(torch.equal(gathered[0], t) for t in gathered), 'parameters differ after attach'
        # ---- two train steps on different shards
        gen = torch.Generator().manual_seed(5)
        X = torch.rand(2, world * 16, 3, 16, 16, generator=gen)
        Y = torch.randint(0, 4, (2, world * 16), generator=gen)
        net.train()
        for s in range(2):
            losses, _ = net.train_step(X[s, rank * 16:(rank + 1) * 16].to(dev), Y[s, rank * 16:(rank + 1) * 16].to(dev))
            assert torch.isfinite(losses['total']).all()
        dist.all_gather(gathered, flat())
        assert all(torch.equal(gathered[0], t) for t in gathered), 'parameters differ after data-parallel steps'
        # BatchNorm running statistics stay per rank (different shards): equalise them for the scoring comparison
        for b in net.buffers():
            dist.broadcast(b.data, src=0)
        pkg.engine.bump_stats()
        # ---- sample-sharded scoring with one final gather
        n = 37
        xs = torch.rand(n, 3, 16, 16, generator=gen).to(dev)
        eps = torch.randn(4, n, 16, generator=gen).to(dev)
        net.eval()

        def scores(lo, hi):
            net.encoder.sampling.injected_eps = eps[:, lo:hi].contiguous()
            with torch.no_grad():
                _, logits, losses, measures = net.evaluate(xs[lo:hi])
                return net.batch_dist_measures(logits, losses, ['elbo', 'iws', 'zdist', 'kl'])

        lo, hi = pkg.distributed.shard_range(n)
        full = pkg.distributed.gather_scores(scores(lo, hi), n, dst=0)
        if rank == 0:
            want = scores(0, n)
            for k in want:
                assert torch.allclose(full[k], want[k].float(), rtol=1e-5, atol=1e-4), k
        else:
            assert full is None
        dist.barrier()
        dist.destroy_process_group()
        out[rank] = 'ok'
    except Exception as e:      # noqa: BLE001 - reported to the parent
        import traceback
        out[rank] = f'{type(e).__name__}: {e}\n{traceback.format_exc()}'


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_data_parallel_and_gather_over_nccl():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: 'ok', 1: 'ok'}, dict(out)
