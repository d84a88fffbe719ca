"""TEST INFRASTRUCTURE: the BASELINE.json configurations as parity cases (shared by tests/golden/make_full_golden.py, which
runs the UNMODIFIED reference on them, and by the CPU / GPU tests).

No multi-megabyte state_dict is committed: every parameter and buffer is filled from a generator seeded by the CRC of its
state_dict key (`fill_state_`), identically in the reference, in the oracle and in the product; inputs come from seeded
generators too.  The fixtures hold the reference's outputs: per-sample / per-class losses, logits, latent tensors, scores,
predictions, a few reconstructed images, and for every parameter gradient its norm and 16 fixed +-1 projections (enough to pin
a full gradient tensor statistically without storing it)."""
import zlib

import numpy as np
import torch

PRIOR = {'init_mean': 1.0, 'learned_means': True, 'var_dim': 'scalar', 'seed': 1}
OPT = {'optim_type': 'adam', 'lr': 1e-3, 'weight_decay': 3e-5, 'grad_clipping': 100}


def _conv(C, K, features='vgg19', upsampler='deconv32', shape=(3, 32, 32), L=16):
    return dict(input_shape=list(shape), num_labels=C, type='cvae', features=features, upsampler=upsampler, encoder=[],
                decoder=[], classifier=[], batch_norm='both', latent_dim=K, latent_sampling=L, test_latent_sampling=L,
                gamma=0, beta=1.0, output_activation='linear', sigma={'value': 1.0, 'learned': True}, optimizer=dict(OPT))


# name -> (constructor keywords, batch).  c1 at its full batch; the conv configurations at a batch the reference's CPU path
# finishes in seconds (BatchNorm then still averages over >= 16 * 17 * 64 values per channel).
CASES = {
    'full_c1': (dict(input_shape=[1, 28, 28], num_labels=10, type='cvae', encoder=[512, 256], decoder=[256, 512], classifier=[],
                     latent_dim=16, latent_sampling=1, test_latent_sampling=1, gamma=0, beta=1.0,
                     output_activation='sigmoid', sigma={'value': 0.1}, optimizer=dict(OPT)), 128),
    'full_c2': (_conv(10, 128), 32),
    'full_c3': (_conv(100, 256), 32),
    'full_c4': (_conv(20, 256, features='resnet18', upsampler='ivgg', shape=(3, 64, 64), L=8), 16),
    # c2 with BatchNorm parameters in the chaotic regime (see fill_state_): reported against loose bounds only
    'full_c2x': (_conv(10, 128), 32),
}
CHAOTIC = {'full_c2x'}
NPROJ = 16


def ctor_kwargs(name):
    import json
    kw = json.loads(json.dumps(CASES[name][0]))
    kw['input_shape'] = tuple(kw['input_shape'])
    kw['prior'] = dict(PRIOR)
    return kw


def _rs(key, salt=0):
    return np.random.RandomState((zlib.crc32(key.encode()) + salt) & 0x7fffffff)


def fill_state_(module, chaotic=False):
    """in place: every entry of module.state_dict() from a generator seeded by its key.

    BatchNorm scale / shift are drawn so that the train-mode network is WELL CONDITIONED (scale in [0.25, 0.75], shift in
    [0.5, 1]: most pre-activations sit on the linear side of the ReLU).  With scale ~ 1, shift ~ 0 a deep random BatchNorm +
    ReLU stack is chaotic: measured on vgg19 in train mode, a 1e-3 perturbation (bf16 rounding of the weights alone, fp32
    everywhere else) grows 1.2x per layer to 3.6e-2 at the features, so no bf16 implementation can be compared at 2e-2 there
    (`chaotic=True` keeps that regime for a loosely-bounded case)."""
    bn_keys = {}
    for mname, m in module.named_modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            bn_keys[mname + '.weight'] = 'scale'
            bn_keys[mname + '.bias'] = 'shift'
    with torch.no_grad():
        for k, v in module.state_dict().items():
            if 'num_batches_tracked' in k or k == 'sigma' or k.endswith('_var_parameter'):
                continue
            rs, shape = _rs(k), tuple(v.shape)
            if k.endswith('running_var'):
                a = rs.uniform(0.5, 1.5, shape)
            elif k.endswith('running_mean'):
                a = rs.normal(0.0, 0.1, shape)
            elif k.endswith('prior.mean'):
                a = rs.normal(0.0, 1.0, shape)
            elif bn_keys.get(k) == 'scale':
                a = rs.uniform(0.5, 1.5, shape) if chaotic else rs.uniform(0.25, 0.75, shape)
            elif bn_keys.get(k) == 'shift':
                a = rs.uniform(-0.1, 0.1, shape) if chaotic else rs.uniform(0.5, 1.0, shape)
            elif v.dim() == 1:                                    # biases
                a = rs.uniform(-0.1, 0.1, shape)
            else:
                fan = v.numel() // shape[0]
                b = (3.0 / fan) ** 0.5
                a = rs.uniform(-b, b, shape)
            v.copy_(torch.from_numpy(np.asarray(a, dtype=np.float32)))
    return module


def inputs(name):
    kw, B = CASES[name]
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7fffffff)
    K, C, L, Lt = kw['latent_dim'], kw['num_labels'], kw['latent_sampling'], kw['test_latent_sampling']
    x = torch.rand(B, *kw['input_shape'], generator=g)
    y = torch.randint(0, C, (B,), generator=g)
    return x, y, torch.randn(L + 1, B, K, generator=g), torch.randn(Lt + 1, B, K, generator=g)


def project(key, g):
    """(norm, NPROJ projections on fixed +-1 vectors) of a gradient tensor"""
    g = np.asarray(g, dtype=np.float64).reshape(-1)
    rs = _rs(key, salt=1)
    out = np.empty(NPROJ)
    for i in range(NPROJ):          # one sign vector at a time: 2.4 M-element tensors stay cheap
        out[i] = np.dot(rs.randint(0, 2, g.size, dtype=np.int8).astype(np.float64) * 2 - 1, g)
    return float(np.linalg.norm(g)), out


def projected_error(key, g, ref_norm, ref_proj):
    """estimate of |g - g_ref| / |g_ref| from the projections: E[(r.(g - g_ref))^2] = |g - g_ref|^2 for Rademacher r"""
    _, p = project(key, g)
    return float(np.sqrt(np.mean((p - ref_proj) ** 2)) / max(ref_norm, 1e-30))


def rank_agreement(got, want, tol):
    """Rank order of a per-sample score (what the ROC tables of OOD / misclassification detection see).
    -> (fraction of ALL sample pairs ordered as in the reference, fraction of the pairs whose reference scores differ by more
    than the margin tol * scale that are ordered as in the reference, fraction of pairs beyond that margin).
    Note that pairs further apart than twice the largest value error keep their order by arithmetic; the informative figures
    are the all-pairs agreement and the agreement beyond a margin SMALLER than the value tolerance."""
    got, want = np.asarray(got, np.float64).reshape(-1), np.asarray(want, np.float64).reshape(-1)
    ok = np.isfinite(want) & np.isfinite(got)
    got, want = got[ok], want[ok]
    scale = max(1e-12, float(np.abs(want).max()))
    iu = np.triu_indices(got.size, 1)
    dw = (want[:, None] - want[None, :])[iu]
    dg = (got[:, None] - got[None, :])[iu]
    same = np.sign(dw) == np.sign(dg)
    clear = np.abs(dw) > tol * scale
    return float(same.mean()) if same.size else 1.0, (float(same[clear].mean()) if clear.any() else 1.0), float(clear.mean()) if clear.size else 0.0


ILL_CONDITIONED = ('nstd', 'IYx', 'mag', 'softiws')      # functions of near-equal exponentials of losses ~ 1e3..1e4


def ill_conditioned(method):
    return method.split('-')[0] in ILL_CONDITIONED


def build_oracle(pkg, name):
    """(product network on the CPU -- parameter containers only, no compute --, OracleNet with the same state).
    torchvision residual features (c4) have no entry in the oracle's layer schema: the oracle then runs a deep copy of the
    torchvision modules themselves in fp32 (what the reference does, conv.py:247-272)."""
    import copy
    from oracle.torch_model import OracleNet, describe_model, describe_seq
    torch.manual_seed(0)
    model = pkg.ClassificationVariationalNetwork(**ctor_kwargs(name))
    fill_state_(model, chaotic=name in CHAOTIC)
    feats = model.features
    try:
        describe_seq(feats)
        cfg, arch = describe_model(model)
        net = OracleNet(cfg, arch)
    except TypeError:
        model.features = None
        cfg, arch = describe_model(model)
        model.features = feats
        arch['features'], arch['features_out'] = [], list(model.encoder.input_shape)
        net = OracleNet(cfg, arch)
        net.features = copy.deepcopy(feats)
        net.encoder = type(net.encoder)(arch, net.K, net.C, int(np.prod(model.encoder.input_shape)))
    net.load_state_dict(model.state_dict())
    return model, net


def _make_bf16_pipeline_(net):
    """In place: every GEMM / convolution weight rounded to bf16, and every tensor that travels between layers -- the output
    of each convolution, linear layer, BatchNorm and activation in the forward pass, and the gradient arriving at it in the
    backward pass -- rounded to bf16, all arithmetic staying fp32.  Nothing here knows the product's kernels: it is what any
    pipeline that stores activations in bf16 and feeds bf16 operands to tensor cores computes."""
    from torch import nn
    rnd = lambda t: t.to(torch.bfloat16).float()
    with torch.no_grad():
        for k, p in net.named_parameters():
            if p.dim() > 1 and 'prior' not in k:
                p.copy_(rnd(p))
    heads = {id(net.encoder.dense_mean), id(net.encoder.dense_log_var)}      # fp32 outputs (they feed the fp32 loss)

    def hook(mod, inp, out):
        out = rnd(out)
        if out.requires_grad:
            out.register_hook(rnd)
        return out

    kinds = (nn.Conv2d, nn.ConvTranspose2d, nn.Linear, nn.BatchNorm2d, nn.ReLU, nn.LeakyReLU, nn.Sigmoid)
    for m in net.modules():
        if isinstance(m, kinds) and id(m) not in heads:
            if isinstance(m, nn.ReLU):
                m.inplace = False
            m.register_forward_hook(hook)
    return net


def oracle_outputs(pkg, name, train_backward=True, bf16_operands=False):
    """the oracle's eval losses / scores / predictions and its training losses + gradients for a case.
    bf16_operands: the same fp32 arithmetic as a GENERIC bf16 tensor-core pipeline would see it (_make_bf16_pipeline_):
    the error this causes against the exact oracle is the floor of any implementation with bf16 operands, which north_star
    mandates."""
    from oracle import elbo_numpy as on
    model, net = build_oracle(pkg, name)
    kw = CASES[name][0]
    x, y, eps_tr, eps_te = inputs(name)
    if bf16_operands:
        _make_bf16_pipeline_(net)
        x = x.to(torch.bfloat16).float()
    n = lambda t: None if t is None else t.detach().numpy()
    prior = on.Prior(n(net.encoder.prior.mean), n(net.encoder.prior._var_parameter), var_dim='scalar', conditional=True)
    skw = dict(sigma_value=float(net.sigma[0]), sigma_is_log=bool(net.arch['sigma']['is_log']),
               sigma_is_rmse=bool(net.arch['sigma']['is_rmse']))
    out = {}
    net.eval()
    with torch.no_grad():
        xr, ye, mu, lv, z, en = net(x, eps_te)
    losses, logits = on.evaluate(n(x), n(xr), n(ye), n(mu), n(lv), n(z), n(en), prior, y=None, training=False,
                                 type='cvae', beta=kw['beta'], gamma=0.0, y_is_decoded=bool(net.arch['y_is_decoded']), **skw)
    out['eval'] = dict(losses=losses, logits=logits, mu=n(mu), x_reco2=n(xr[:2, :2]))
    if train_backward:
        net.train()
        tl, (xr, ye, mu, lv, z) = net.train_losses(x, y, eps_tr, beta=kw['beta'], gamma=0.0)
        tl['total'].mean().backward()
        out['train'] = dict(losses={k: n(v) for k, v in tl.items()}, mu=n(mu), x_reco2=n(xr[:2, :2]),
                            grads={k: n(p.grad) for k, p in net.named_parameters() if p.grad is not None})
    return out
