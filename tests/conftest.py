import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    # a fresh checkout has no lib/libmsda_b200.so (build artefacts are git-ignored): build it once, here, exactly as
    # __graft_entry__.build() does, so the ABI tests exercise the real library. If nvcc is missing the product's own
    # loud failure ("... is not built ...") is what the tests then report.
    try:
        from vit_adapter_b200 import build as _build
        if _build.is_stale():
            _build.build()
    except Exception as exc:  # pragma: no cover
        sys.stderr.write('conftest: could not build the CUDA library: %r\n' % (exc,))


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """npz fixture -> dict of torch tensors (made by tests/golden/make_golden.py from the real reference)."""
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    return {k: torch.from_numpy(z[k]) for k in z.files}


OP_CASES = ['op_kat_seed3', 'op_inj_edges', 'op_ext_edges', 'op_odd_d5', 'op_d32_edges', 'op_d64_edges']


def make_inputs(N, M, D, Lq, shapes, P, seed, dist='uniform', dtype=torch.float32):
    """Seeded synthetic operator inputs on the CPU generator (as detection/ops/test.py does).

    dist: 'uniform'  loc ~ U(0,1), aw ~ U+1e-5 normalised                 (ops/test.py:28-33)
          'edges'    loc ~ U(-0.1,1.1) + exact texel centres / 0 / 1       (every validity branch)
          'adapter'  reference-point grid + init-bias ring + N(0,1px) noise, softmax(N(0,1)) weights
    """
    g = torch.Generator().manual_seed(seed)
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    L = shapes_t.shape[0]
    S = int(shapes_t.prod(1).sum())
    lsi = torch.cat((shapes_t.new_zeros((1,)), shapes_t.prod(1).cumsum(0)[:-1]))
    value = torch.randn(N, S, M, D, generator=g)
    if dist == 'uniform':
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g)
        aw = torch.rand(N, Lq, M, L, P, generator=g) + 1e-5
        aw = aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    elif dist == 'edges':
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * 1.2 - 0.1
        flat = loc.view(-1, 2)
        W0 = float(shapes_t[0, 1])
        n7 = flat[0::7].shape[0]
        flat[0::7] = (torch.randint(0, int(W0), (n7, 2), generator=g).float() + 0.5) / W0
        flat[1::11] = 0.0
        flat[2::13] = 1.0
        aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    elif dist == 'adapter':
        import math
        side = int(round(math.sqrt(Lq)))
        if side * side == Lq:
            ys, xs = torch.meshgrid((torch.arange(side) + 0.5) / side, (torch.arange(side) + 0.5) / side, indexing='ij')
            ref = torch.stack([xs.reshape(-1), ys.reshape(-1)], -1)
        else:
            ref = torch.rand(Lq, 2, generator=g)
        theta = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
        ray = torch.stack([theta.cos(), theta.sin()], -1)
        ray = ray / ray.abs().max(-1, keepdim=True)[0]
        off = ray.view(1, 1, M, 1, 1, 2) * torch.arange(1, P + 1).view(1, 1, 1, 1, P, 1)
        off = off + torch.randn(N, Lq, M, L, P, 2, generator=g)
        wh = torch.stack([shapes_t[:, 1], shapes_t[:, 0]], -1).float()
        loc = ref.view(1, Lq, 1, 1, 1, 2) + off / wh.view(1, 1, 1, L, 1, 2)
        aw = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    else:
        raise ValueError(dist)
    grad_out = torch.randn(N, Lq, M * D, generator=g)
    return dict(value=value.to(dtype), shapes=shapes_t, lsi=lsi, loc=loc.to(dtype).contiguous(),
                aw=aw.to(dtype).contiguous(), grad_out=grad_out.to(dtype))
