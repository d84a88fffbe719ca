#!/usr/bin/env python
"""Benchmark of the joint-VAE hot path on B200 (BASELINE.json metric: train images/s on the CIFAR-10-shaped conv
joint-VAE, K=128, L=16, C=10, batch 512 per GPU, bf16 GEMMs; plus OOD-scoring samples/s and the fused-ELBO HBM GB/s).

  python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                    the reference's algorithm (oracle port) on the host cores

A "step" is one full optimisation step of the path: zero_grad, features/encoder/sampler/decoder/classifier forward,
fused ELBO forward + backward, network backward, global-norm clip, Adam.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: conv joint-VAE on synthetic 3x32x32, VGG-style encoder, K=128, L=16, C=10, B=512/GPU
    'c2': dict(ctor=dict(input_shape=(3, 32, 32), num_labels=10, type='cvae', features='vgg19', upsampler='deconv32',
                         encoder=[], decoder=[], classifier=[], batch_norm='both', latent_dim=128, latent_sampling=16,
                         test_latent_sampling=16, gamma=0, beta=1.0, output_activation='linear',
                         sigma={'value': 1.0, 'learned': True},
                         optimizer={'optim_type': 'adam', 'lr': 1e-3, 'weight_decay': 3e-5, 'grad_clipping': 100}),
               batch=512, fwd_gflop_per_img=2.903,
               name='conv joint-VAE 3x32x32 vgg19+deconv32 BN K=128 L=16 C=10 B=512/GPU'),
    # BASELINE.json configs[2]: CIFAR-100 shape, wider per-class contraction (parity-test case; optional bench workload)
    'c3': dict(ctor=dict(input_shape=(3, 32, 32), num_labels=100, type='cvae', features='vgg19', upsampler='deconv32',
                         encoder=[], decoder=[], classifier=[], batch_norm='both', latent_dim=256, latent_sampling=16,
                         test_latent_sampling=16, gamma=0, beta=1.0, output_activation='linear',
                         sigma={'value': 1.0, 'learned': True},
                         optimizer={'optim_type': 'adam', 'lr': 1e-3, 'weight_decay': 3e-5, 'grad_clipping': 100}),
               batch=512, fwd_gflop_per_img=2.921,
               name='conv joint-VAE 3x32x32 vgg19+deconv32 BN K=256 L=16 C=100 B=512/GPU'),
    # BASELINE.json configs[3]: ResNet-style, imagenet20-subset shape (parity-test case; optional bench workload)
    'c4': dict(ctor=dict(input_shape=(3, 64, 64), num_labels=20, type='cvae', features='resnet18', upsampler='ivgg',
                         encoder=[], decoder=[], classifier=[], batch_norm='both', latent_dim=256, latent_sampling=8,
                         test_latent_sampling=8, gamma=0, beta=1.0, output_activation='linear',
                         sigma={'value': 1.0, 'learned': True},
                         optimizer={'optim_type': 'adam', 'lr': 1e-3, 'weight_decay': 3e-5, 'grad_clipping': 100}),
               batch=128, fwd_gflop_per_img=None,
               name='ResNet joint-VAE 3x64x64 resnet18+ivgg BN K=256 L=8 C=20 B=128/GPU'),
    # BASELINE.json configs[0]: the reference's CPU-runnable case
    'c1': dict(ctor=dict(input_shape=(1, 28, 28), num_labels=10, type='cvae', encoder=[512, 256], decoder=[256, 512],
                         classifier=[], latent_dim=16, latent_sampling=1, test_latent_sampling=1, gamma=0, beta=1.0,
                         output_activation='sigmoid', sigma={'value': 0.1},
                         optimizer={'optim_type': 'adam', 'lr': 1e-3, 'weight_decay': 3e-5, 'grad_clipping': 100}),
               batch=128, fwd_gflop_per_img=0.0032, name='MLP joint-VAE 1x28x28 K=16 L=1 C=10 B=128'),
}
PRIOR = {'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 1}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops_sustained', 1400.0), 'measured'
    return 6650.0, 1590.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)"""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


def make_ctor(wl):
    kw = json.loads(json.dumps(wl['ctor']))
    kw['input_shape'] = tuple(kw['input_shape'])
    kw['prior'] = dict(PRIOR)
    return kw


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_run(wl, steps, warmup, sample_batch, threads=None):
    """The reference's algorithm for the same step (oracle/torch_model.py, a PyTorch fp32 restatement of cvae.py's train
    step pinned against the unmodified reference) on the host cores."""
    import torch
    import __graft_entry__ as g
    from oracle.torch_model import OracleNet, describe_model, train_step
    if threads:
        torch.set_num_threads(threads)
    pkg = g.load_package()
    torch.manual_seed(0)
    model = pkg.ClassificationVariationalNetwork(**make_ctor(wl))          # layer containers only (CPU, no compute)
    cfg, arch = describe_model(model)
    net = OracleNet(cfg, arch)
    net.load_state_dict(model.state_dict())
    net.train()
    oc = wl['ctor']['optimizer']
    opt = torch.optim.Adam(net.parameters(), lr=oc['lr'], weight_decay=oc['weight_decay'])
    B, L, K, C = sample_batch, wl['ctor']['latent_sampling'], wl['ctor']['latent_dim'], wl['ctor']['num_labels']
    gen = torch.Generator().manual_seed(0)
    x = torch.rand(B, *wl['ctor']['input_shape'], generator=gen)
    y = torch.randint(0, C, (B,), generator=gen)
    times = []
    for i in range(warmup + steps):
        eps = torch.randn(L + 1, B, K, generator=gen)
        t0 = time.perf_counter()
        train_step(net, opt, x, y, eps, beta=wl['ctor']['beta'], gamma=wl['ctor']['gamma'] or 0.0, clip=oc['grad_clipping'])
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return B / dt, dt, torch.get_num_threads()


def run_reference(args, wl):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sb = args.cpu_batch
    v, dt, cores = cpu_reference_run(wl, args.steps, args.warmup, sb)
    line = {'impl': 'reference', 'metric': 'train_images_per_sec', 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'sample_batch': sb},
            'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                             'sample': f'{args.steps} full train steps at batch {sb} (CPU time is linear in the batch)'},
            'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ native arm (GPU)
def run_native(args, wl):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    pkg = g.build()
    nat = pkg._native
    if args.conv:
        pkg.engine.POLICY['conv'] = args.conv
    if args.linear:
        pkg.engine.POLICY['linear'] = args.linear
    torch.manual_seed(0)
    net = pkg.ClassificationVariationalNetwork(**make_ctor(wl)).to(dev)
    if world > 1:
        pkg.distributed.attach(net, bf16_bucket=True)
    net.train()
    B = args.batch or wl['batch']
    C = wl['ctor']['num_labels']
    shape = tuple(wl['ctor']['input_shape'])
    gen = torch.Generator().manual_seed(1 + rank)
    npool = 4
    xs_h = [torch.rand(B, *shape, generator=gen).pin_memory() for _ in range(npool)]
    ys_h = [torch.randint(0, C, (B,), generator=gen).pin_memory() for _ in range(npool)]
    xs = [t.to(dev) for t in xs_h]
    ys = [t.to(dev) for t in ys_h]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(i)
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    step_dev = lambda i: net.train_step(xs[i % npool], ys[i % npool])

    def step_e2e(i):
        x = xs_h[i % npool].to(dev, non_blocking=True)
        y = ys_h[i % npool].to(dev, non_blocking=True)
        losses, _ = net.train_step(x, y)
        return float(losses['total'].mean().item())          # device -> host read of the step's result

    for i in range(args.warmup):
        step_dev(i)
    # ---- device-resident timing (value), with live per-launch timing of the fused ELBO kernels
    nat.PROFILE = {'elbo_train_fwd': [], 'elbo_train_bwd': []}
    clocks = ClockSampler(local)
    n0 = nat.launch_count()
    ms = timed(step_dev, args.steps)
    launches = nat.launch_count() - n0
    clk = clocks.stop()
    prof = {k: [a.elapsed_time(b) for a, b in v] for k, v in nat.PROFILE.items()}
    nat.PROFILE = None
    # ---- end-to-end timing through the public API with host buffers
    step_e2e(0)
    ms_e2e = timed(step_e2e, args.steps)
    # ---- the same, fed by the uint8 input pipeline (utils/batch_loader.py): dataset in pinned HOST memory, per step one
    # uint8 batch copy + gather / flip / random-crop / ToTensor kernel on a side stream, train step, loss read back
    ms_loader = None
    if len(shape) == 3:
        from jointvae_b200.utils.batch_loader import DeviceBatchLoader
        u8 = torch.randint(0, 256, ((args.steps + 2) * B, shape[1], shape[2], shape[0]), dtype=torch.uint8, generator=gen)
        tg = torch.randint(0, C, (u8.shape[0],), generator=gen)
        loader = DeviceBatchLoader(u8, tg, B, device=dev, data_augmentation=['flip', 'crop'], resident=False, seed=rank)
        it = iter(loader)
        step_loader = lambda i: float(net.train_step(*next(it))[0]['total'].mean().item())
        step_loader(0)
        ms_loader = timed(step_loader, args.steps)

    # ---- OOD scoring throughput (per-class evaluate + scores + predictions), device resident
    n_methods = len(net.ood_methods)
    net.eval()
    with torch.no_grad():
        def score(i):
            _, logits, losses, _ = net.evaluate(xs[i % npool])
            net.batch_dist_measures(logits, losses, [m for m in net.ood_methods])
            net.predict_after_evaluate(logits, losses, method=net.predict_methods[0])
        for i in range(6):          # the caching allocator re-sizes its pools when the step shape changes: settle first
            score(i)
        timed(score, 3)
        nat.PROFILE = {'elbo_eval_fwd': []}
        ms_score = min(timed(score, max(3, args.steps)), timed(score, max(3, args.steps)))     # best of two runs
        prof['elbo_eval_fwd'] = [a.elapsed_time(b) for a, b in nat.PROFILE['elbo_eval_fwd']]
        nat.PROFILE = None

    # ---- BASELINE configs[4]: OOD-scoring sweep over the class count (C = 10 / 100 / 1000, K = 256, L = 16), sample-sharded:
    # every rank scores its own shard (no data-path collective); rate = samples of all ranks / max-over-ranks time
    sweep = []
    if args.sweep:
        for C_ in (10, 100, 1000):
            kw = make_ctor(wl)
            kw.update(num_labels=C_, latent_dim=256)
            torch.manual_seed(0)
            m = pkg.ClassificationVariationalNetwork(**kw).to(dev)
            m.eval()
            xs_ = [torch.rand(B, *shape, device=dev) for _ in range(2)]
            with torch.no_grad():
                def sc(i, m=m, xs_=xs_):
                    _, lg, ls, _ = m.evaluate(xs_[i % 2])
                    m.batch_dist_measures(lg, ls, list(m.ood_methods))
                    m.predict_after_evaluate(lg, ls, method=m.predict_methods[0])
                for i in range(6):
                    sc(i)
                timed(sc, 5)               # settles the caching allocator for this model's shapes
                nat.PROFILE = {'elbo_eval_fwd': []}
                ms_c = min(timed(sc, 5), timed(sc, 5), timed(sc, 5))      # best of three 5-batch runs
                ev = [a.elapsed_time(b) for a, b in nat.PROFILE['elbo_eval_fwd']]
                nat.PROFILE = None
            sweep.append({'C': C_, 'K': 256, 'L': wl['ctor']['test_latent_sampling'], 'samples_per_s': world * B * 5 / (ms_c * 1e-3),
                          'eval_kernel_us': sum(ev) / max(1, len(ev)) * 1e3})
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, tf, which = peaks()
    L, K, D = wl['ctor']['latent_sampling'], wl['ctor']['latent_dim'], int(torch.tensor(shape).prod())
    has_xr = True
    # algorithmic bytes of the fused ELBO train forward per sample: x f32 + L reconstructions (bf16)
    # + mu/log_var f32 + label + 8 outputs (logits only when a classifier exists: gamma=0 here)
    bytes_fwd = B * (D * 4 + L * D * 2 + 2 * K * 4 + 8 + 8 * 4)
    t_fwd = sum(prof['elbo_train_fwd']) / max(1, len(prof['elbo_train_fwd'])) * 1e-3
    achieved = bytes_fwd / t_fwd / 1e9 if t_fwd > 0 else None
    value = world * B * args.steps / (ms * 1e-3)
    e2e = world * B * args.steps / (ms_e2e * 1e-3)
    flops = 3 * (wl['fwd_gflop_per_img'] or 0.0) * 1e9 * B * args.steps / (ms * 1e-3)
    line = {
        'metric': 'train_images_per_sec', 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': wl['name'], 'global_batch': world * B, 'parallelism': f'dp{world}',
                   'l2': 'per-step working set (activations, 53 MB x_reco alone) exceeds L2; 4 rotating input batches',
                   'backend': dict(pkg.engine.POLICY)},
        'clocks': clk,
        'e2e': {'value': e2e, 'unit': 'images/s', 'h2d_bytes_per_step': B * D * 4 + B * 8, 'd2h_bytes_per_step': 4,
                'ms_per_step': ms_e2e / args.steps},
        'e2e_uint8_loader': None if ms_loader is None else {
            'value': world * B * args.steps / (ms_loader * 1e-3), 'unit': 'images/s', 'h2d_bytes_per_step': B * D + B * 25,
            'd2h_bytes_per_step': 4, 'note': 'uint8 dataset in pinned host memory, flip + crop + ToTensor on the device'},
        'gpu_launches': int(launches),
        'roofline': {'kernel': 'elbo_train_fwd_kernel (fused prior/ELBO forward)', 'bound': 'hbm', 'achieved': achieved,
                     'peak': hbm, 'unit': 'GB/s', 'frac': (achieved / hbm) if achieved else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
                     # of this kernel at this workload (profiles/r01_elbo_train_fwd_ncu_full.txt): 57.23 MB + 0.36 MB
                     'traffic': 57.59e6 if (args.workload == 'c2' and B == 512) else None,
                     'peak_source': which + ' (MEASURED_PEAKS.json hbm_gbs)', 'bytes_per_launch': bytes_fwd,
                     'us_per_launch': t_fwd * 1e6,
                     'bwd_us_per_launch': sum(prof['elbo_train_bwd']) / max(1, len(prof['elbo_train_bwd'])) * 1e3,
                     'eval_us_per_launch': sum(prof['elbo_eval_fwd']) / max(1, len(prof['elbo_eval_fwd'])) * 1e3},
        'gemm_roofline': None if not wl['fwd_gflop_per_img'] else {
            'bound': 'tensor', 'achieved': flops / 1e12, 'peak': tf, 'unit': 'TFLOP/s', 'frac': flops / 1e12 / tf,
            'note': 'whole step: 3 x forward GEMM/conv FLOPs / step time'},
        'scoring': {'value': world * B * max(3, args.steps) / (ms_score * 1e-3), 'unit': 'samples/s',
                    'methods': n_methods},
        'scoring_sweep': sweep,
    }
    if world == 1 and not args.no_cpu:
        v, dt, cores = cpu_reference_run(wl, 2, 1, args.cpu_batch)
        line['cpu_baseline'] = {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                                'sample': f'2 full train steps at batch {args.cpu_batch} after 1 warm-up, oracle/torch_model.py fp32'}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0)
    ap.add_argument('--cpu-batch', type=int, default=32)
    ap.add_argument('--conv', default='', choices=['', 'native', 'library'])
    ap.add_argument('--linear', default='', choices=['', 'native', 'library'])
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-sweep', dest='sweep', action='store_false', help='skip the C = 10/100/1000 scoring sweep')
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        run_native(args, wl)


if __name__ == '__main__':
    main()
