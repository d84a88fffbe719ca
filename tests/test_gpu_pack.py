"""The native weight re-pack (csrc/pack.cu: every bf16 operand arrangement of a stack in one launch) against its torch
definition (conv_engine.run_pack_spec, itself pinned to the torch packing code on CPU by tests/test_conv_engine_cpu.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

CASES = [
    ('input', 'vgg19', (3, 32, 32), True, None),
    ('input', '16x3+1:2-8x5+2', (5, 9, 9), False, None),
    ('input', 'resnet18', (3, 64, 64), True, None),
    ('output', 'deconv32', (128, 1, 1), True, 'linear'),
    ('output', 'ivgg', (16, 4, 4), True, 'linear'),
    ('output', '[x4+1]8x4+1:2-!3x3+1', (6, 3, 3), False, 'linear'),
]


@pytest.mark.parametrize('where,spec,shape,bn,out_act', CASES)
def test_pack_kernel_matches_definition(pkg, where, spec, shape, bn, out_act):
    from jointvae_b200 import conv_engine as ce
    torch.manual_seed(0)
    kw = dict(output_activation=out_act) if where == 'output' else {}
    seq = pkg.module.vae_layers.build_de_conv_layers(shape, spec, batch_norm=bn, where=where, **kw).to(DEV)
    stack = ce.ConvStack(list(seq), shape, where == 'output')
    n0 = pkg._native.launch_count()
    stack.plan.ensure()
    assert pkg._native.launch_count() == n0 + 1, 'the whole stack is one launch'
    torch.cuda.synchronize()
    n = 0
    for stp in stack.plan.steps:
        pk = stp._pack()
        for sp in stp.pack_specs():
            src = stp.conv.weight if sp['src'] == 'w' else stp.conv.bias
            want = stp.pack_view(sp, ce.run_pack_spec(sp, src))
            got = pk[sp['name']] if sp['index'] is None else pk[sp['name']][sp['index']]
            assert got.shape == want.shape and got.dtype == want.dtype
            assert torch.equal(got, want), (type(stp.conv).__name__, sp['name'], sp['index'])
            n += 1
    assert n >= 2
    # a raw in-place parameter update + epoch bump (what Optimizer.step does) refreshes every arrangement, again in one launch
    with torch.no_grad():
        for stp in stack.plan.steps:
            stp.conv.weight.data.mul_(1.5)
    pkg.engine.bump_params()
    n0 = pkg._native.launch_count()
    stack.plan.ensure()
    assert pkg._native.launch_count() == n0 + 1
    stp = stack.plan.steps[-1]
    sp = stp.pack_specs()[0]
    want = stp.pack_view(sp, ce.run_pack_spec(sp, stp.conv.weight))
    got = stp._pk[sp['name']] if sp['index'] is None else stp._pk[sp['name']][sp['index']]
    assert torch.equal(got, want)
