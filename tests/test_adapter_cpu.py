"""Adapter-level host logic without a GPU: deform_inputs and parameter naming vs the real reference
(golden: tests/golden/adapter_block.npz, made from detection/mmdet_custom/.../adapter_modules.py)."""
import torch

from conftest import load_golden

from vit_adapter_b200.adapter import (InteractionBlock, InteractionBlockWithCls, InteractionBlockWithText, SpatialPriorModule, deform_inputs,
                                      get_reference_points)


def test_deform_inputs_match_reference():
    g = load_golden('adapter_block')
    dim, heads, H, W, N = [int(v) for v in g['cfg']]
    d1, d2 = deform_inputs(torch.zeros(N, 3, H, W))
    torch.testing.assert_close(d1[0], g['ref1'], rtol=0, atol=0)
    assert torch.equal(d1[1], g['shapes1']) and torch.equal(d1[2], g['lsi1'])
    torch.testing.assert_close(d2[0], g['ref2'], rtol=0, atol=0)
    assert torch.equal(d2[1], g['shapes2']) and torch.equal(d2[2], g['lsi2'])
    assert d1[1].dtype == torch.int64 and d1[0].dtype == torch.float32
    # memoised: the same tensors come back (stable storage => MSDeformAttn's cached shape check, CUDA graphs)
    e1, e2 = deform_inputs(torch.zeros(1, 3, H, W))
    assert e1[1].data_ptr() == d1[1].data_ptr() and e2[0].data_ptr() == d2[0].data_ptr()


def test_reference_points_are_cell_centres():
    r = get_reference_points([(2, 4)], 'cpu')
    assert r.shape == (1, 8, 1, 2)
    torch.testing.assert_close(r[0, :, 0, 0], torch.tensor([0.125, 0.375, 0.625, 0.875] * 2))
    torch.testing.assert_close(r[0, :, 0, 1], torch.tensor([0.25] * 4 + [0.75] * 4))


def test_interaction_block_state_dict_matches_reference():
    g = load_golden('adapter_block')
    dim, heads, H, W, N = [int(v) for v in g['cfg']]
    blk = InteractionBlock(dim=dim, num_heads=heads, n_points=4, init_values=0., deform_ratio=float(g['ratio']),
                           extra_extractor=True, with_cffn=True, cffn_ratio=0.25)
    ref = {k[3:]: v for k, v in g.items() if k.startswith('sd.')}
    sd = blk.state_dict()
    assert sorted(sd.keys()) == sorted(ref.keys())
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    blk.double().load_state_dict(ref, strict=True)
    blk2 = InteractionBlockWithCls(dim=dim, num_heads=heads, deform_ratio=float(g['ratio']), extra_extractor=True)
    assert sorted(blk2.state_dict().keys()) == sorted(ref.keys())
    assert float(blk2.injector.gamma.abs().max()) == 0.0  # init_values = 0 => injector is the identity at init
    # the wsdm2023 copy (text tokens through the blocks) has the same parameters: its golden's keys load strictly
    gt = load_golden('adapter_block_text')
    blk3 = InteractionBlockWithText(dim=dim, num_heads=heads, deform_ratio=float(gt['ratio']), extra_extractor=True).double()
    blk3.load_state_dict({k[3:]: v for k, v in gt.items() if k.startswith('sd.')}, strict=True)


def test_spatial_prior_module_shapes_and_keys():
    spm = SpatialPriorModule(inplanes=8, embed_dim=16, norm_layer=torch.nn.BatchNorm2d)
    c1, c2, c3, c4 = spm(torch.randn(2, 3, 64, 96))
    assert c1.shape == (2, 16, 16, 24) and c2.shape == (2, 8 * 12, 16) and c3.shape == (2, 4 * 6, 16) and c4.shape == (2, 2 * 3, 16)
    keys = set(spm.state_dict().keys())
    for k in ('stem.0.weight', 'stem.1.weight', 'stem.3.weight', 'stem.6.weight', 'stem.7.running_mean', 'conv2.0.weight',
              'conv3.1.bias', 'conv4.0.weight', 'fc1.weight', 'fc4.bias'):
        assert k in keys, k


def test_deform_inputs_cache_survives_inference_mode_and_caller_edits():
    """the first call for a resolution may happen under inference_mode (a sanity validation pass); the memoised
    tensors must still be usable in a later training forward (not inference tensors), and a caller editing the returned
    lists must not corrupt the cache."""
    from vit_adapter_b200.adapter import adapter_modules as am
    am._DEFORM_CACHE.clear()
    with torch.inference_mode():
        d1, d2 = deform_inputs(torch.zeros(1, 3, 96, 64))
    assert not d1[0].is_inference() and not d2[1].is_inference()
    w = torch.ones(1, requires_grad=True)
    (d1[0] * w).sum().backward()                     # saving the memoised tensor for backward works
    assert w.grad is not None
    d1[0] = None                                     # caller-side edit of ITS list
    d2.append('junk')
    e1, e2 = deform_inputs(torch.zeros(2, 3, 96, 64))
    assert e1[0] is not None and len(e2) == 3
    f1, _ = deform_inputs(torch.zeros(2, 3, 96, 64))
    assert f1[0] is e1[0] and f1 is not e1           # same memoised tensors, fresh containers
