#!/usr/bin/env python
"""Per-layer timing of the c2 decoder (deconv32 on (L+1)*B = 8704 samples) and encoder (vgg19, B = 512) through
conv_engine, forward and backward, with CUDA events.  Run on the GPU box:  python tools/bench_layers.py [imager|features]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g

pkg = g.build()
from jointvae_b200 import conv_engine as ce
nat = pkg._native
which = sys.argv[1] if len(sys.argv) > 1 else 'imager'
dev = 'cuda:0'
torch.manual_seed(0)
if which == 'imager':
    seq = pkg.module.vae_layers.build_de_conv_layers((128, 1, 1), 'deconv32', batch_norm=True, where='output',
                                                     output_activation='linear').to(dev).train()
    x = torch.randn(8704, 128, 1, 1, device=dev, requires_grad=True)
    image_out = True
else:
    seq = pkg.module.vae_layers.build_de_conv_layers((3, 32, 32), 'vgg19', batch_norm=True, where='input').to(dev).train()
    x = torch.rand(512, 3, 32, 32, device=dev)
    image_out = False

# wrap the native entry points with CUDA-event timers
timers = []
def wrap(name):
    fn = getattr(ce.K, name)
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record()
        desc = ''
        if name == 'gather':
            desc = f'in{tuple(a[0].shape)} Cin={a[1]} taps={len(a[4][0])} s={a[5]} grid={a[6]}x{a[7]} Cout={a[9]}'
        elif name == 'subpixel':
            desc = f'in{tuple(a[0].shape)} Cin={a[1]} taps={a[4][5]} phases={len(a[4][0])} grid={a[5]}x{a[6]} Cout={a[8]}' + ('' if r else '  NOT COVERED')
        elif name == 'wgrad':
            desc = f'g{tuple(a[0].shape)} x{tuple(a[2].shape)} taps={len(a[4][0])} s={a[5]}'
        elif name == 'gemm':
            desc = f'mode={a[0]} M={a[1]} N={a[2]} K={a[3]}'
        timers.append((name, desc, e0, e1))
        return r
    return staticmethod(w)
for n in ('gather', 'subpixel', 'wgrad', 'gemm', 'bn_apply_fwd', 'bn_bwd', 'bn_stats', 'act_bwd', 'maxpool_fwd', 'maxpool_bwd'):
    setattr(ce.K, n, wrap(n))

for it in range(3):
    timers.clear()
    out = ce.run(list(seq), x, image_out=image_out)
    gout = torch.randn_like(out)
    torch.cuda.synchronize()
    tf = time.perf_counter()
    out.backward(gout)
    torch.cuda.synchronize()
tot = 0.0
for name, desc, e0, e1 in timers:
    ms = e0.elapsed_time(e1)
    tot += ms
    print(f'{ms * 1e3:9.1f} us  {name:14s} {desc}')
print(f'total {tot:.2f} ms (sum of kernel-side spans)')
