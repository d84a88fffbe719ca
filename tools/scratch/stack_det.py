"""Determinism probe of the conv stacks: every native 'gather' launch is repeated with the same operands and compared bitwise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import __graft_entry__ as g
pkg = g.build()
from jointvae_b200 import conv_engine as ce
dev = 'cuda:0'
torch.manual_seed(0)
bad = []
orig = ce.K.gather
def chk(*a, **k):
    out = a[8]
    stats = a[14]
    s0 = stats.clone() if stats is not None else None
    r = orig(*a, **k)
    ref = out.clone()
    for rep in range(6):
        if stats is not None: stats.copy_(s0)
        pkg._native.lib().jvae_probe_poison([0x7fc07fc0, 0x3c003c00, 0x7f807f80, 0, 0xffffffff, 0x42004200][rep], pkg._native.stream())
        orig(*a, **k)
        if not torch.equal(out, ref):
            d = (out.float() - ref.float()).abs()
            bad.append((tuple(a[0].shape), a[1], len(a[4][0]), a[5], a[6], a[7], a[9], a[11], a[12], rep, float(d.max()), int((d > 0).sum())))
            break
    return r
ce.K.gather = staticmethod(chk)
which = sys.argv[1] if len(sys.argv) > 1 else 'small'
if which == 'small':
    up = pkg.module.vae_layers.build_de_conv_layers((16, 1, 1), '[x3+1]16x4+0-16-8:2++1-8:2++1-!3x3+1', batch_norm=True, where='output', output_activation='linear').to(dev).train()
    ft = pkg.module.vae_layers.build_de_conv_layers((3, 16, 16), '[x3+1]8-8-M-16:2-16', batch_norm=True, where='input').to(dev).train()
    xs = [(up, torch.randn(128, 16, 1, 1, device=dev, requires_grad=True), True), (ft, torch.rand(32, 3, 16, 16, device=dev, requires_grad=True), False)]
else:
    up = pkg.module.vae_layers.build_de_conv_layers((128, 1, 1), 'deconv32', batch_norm=True, where='output', output_activation='linear').to(dev).train()
    xs = [(up, torch.randn(8704, 128, 1, 1, device=dev, requires_grad=True), True)]
for seq, x, io in xs:
    for it in range(2):
        out = ce.run(list(seq), x, image_out=io)
        out.backward(torch.randn_like(out))
        torch.cuda.synchronize()
print('nondeterministic launches:', len(bad))
for b in bad: print(b)
