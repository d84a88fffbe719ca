import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
pkg = g.build()
from jointvae_b200 import conv_engine as ce
dev = 'cuda:0'
torch.manual_seed(0)
bl = pkg.module.vae_layers.build_de_conv_layers
def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm()), float((a != b).float().mean())
for name, seq, x in [
    ('vgg19 features', bl((3, 32, 32), 'vgg19', batch_norm=True, where='input').to(dev).train(), torch.rand(512, 3, 32, 32, device=dev)),
    ('deconv32 imager', bl((128, 1, 1), 'deconv32', batch_norm=True, where='output', output_activation='linear').to(dev).train(),
     torch.randn(8704, 128, 1, 1, device=dev)),
    ('deconv32 imager no BN', bl((128, 1, 1), 'deconv32', batch_norm=False, where='output', output_activation='linear').to(dev).train(),
     torch.randn(2048, 128, 1, 1, device=dev)),
]:
    with torch.no_grad():
        a = ce.run(list(seq), x, image_out=('imager' in name))
        b = ce.run(list(seq), x, image_out=('imager' in name))
        print(name, 'rel diff %.3e, fraction of differing elements %.3e' % rel(a, b), flush=True)
        # layer by layer
        st = ce._stacks[list(ce._stacks)[-1]]
        t1 = ce.K.to_nhwc(x, ce.r8(x.shape[1])); t2 = t1.clone()
        for i, s in enumerate(st.steps):
            t1 = s.forward(t1, {}, True); t2 = s.forward(t2, {}, True)
            r = rel(t1[..., :s.out_shape[0]], t2[..., :s.out_shape[0]])
            print('   step', i, type(s).__name__, s.out_shape, 'rel %.3e frac %.3e' % r, flush=True)
