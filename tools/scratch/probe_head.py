import sys, os, copy
sys.path.insert(0, '.'); sys.path.insert(0, './tests')
import numpy as np, torch
from torch import nn
import __graft_entry__ as g
pkg = g.build()
import full_cases as fc
from jointvae_b200 import conv_engine as ce
DEV = 'cuda:0'
torch.manual_seed(0)
seq = pkg.module.vae_layers.build_de_conv_layers((128, 1, 1), 'deconv32', batch_norm=True, where='output', output_activation='linear')


class W(nn.Module):
    def __init__(s, q):
        super().__init__()
        s.imager = q


fc.fill_state_(W(seq))
seq = seq.to(DEV).train()
N = 544
x = torch.randn(N, 128, 1, 1, device=DEV).to(torch.bfloat16).float()
tgt = torch.rand(N, 3, 32, 32, device=DEV)


def rnd(t):
    return t.to(torch.bfloat16).float()


def run_ref(mods, hooks):
    mods = copy.deepcopy(mods)
    if hooks:
        with torch.no_grad():
            for p in mods.parameters():
                if p.dim() > 1:
                    p.copy_(rnd(p))

        def hook(m, i, o):
            o = rnd(o)
            if o.requires_grad:
                o.register_hook(rnd)
            return o
        for m in mods.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.BatchNorm2d, nn.ReLU)):
                if isinstance(m, nn.ReLU):
                    m.inplace = False
                m.register_forward_hook(hook)
    out = mods(x)
    loss = ((out - tgt) ** 2).sum() / N
    loss.backward()
    return out.detach(), {k: p.grad.clone() for k, p in mods.named_parameters()}


o_ref, g_ref = run_ref(seq, False)
o_bf, g_bf = run_ref(seq, True)
for sep in ('1', '0'):
    os.environ['JVAE_CONV_SEPARABLE'] = sep
    ce._stacks.clear()
    mods = copy.deepcopy(seq)
    out = ce.run(list(mods), x.clone(), image_out=True)
    go = (2 * (out.float() - tgt) / N)
    out.backward(go.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
    print('separable', sep, 'fwd err', float((out.float() - o_ref).norm() / o_ref.norm()), 'generic', float((o_bf - o_ref).norm() / o_ref.norm()))
    for k, p in mods.named_parameters():
        nr = g_ref[k].norm()
        if nr < 1e-4:
            continue
        print('   %-12s product %.4f  generic-bf16 %.4f' % (k, float((p.grad - g_ref[k]).norm() / nr), float((g_bf[k] - g_ref[k]).norm() / nr)))
