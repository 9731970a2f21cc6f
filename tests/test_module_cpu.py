"""Host-side mirror of the reference module API (no GPU): constructor contract, parameter names,
initialisation identical to the reference's _reset_parameters (golden from the real reference)."""
import warnings

import pytest
import torch

from conftest import load_golden

from vit_adapter_b200 import MSDeformAttn


def test_state_dict_keys_and_shapes_match_reference():
    g = load_golden('module_l3')
    d_model, L, M, P = [int(x) for x in g['cfg']]
    m = MSDeformAttn(d_model, L, M, P, float(g['ratio']))
    sd = m.state_dict()
    ref_keys = sorted(k[3:] for k in g if k.startswith('sd.'))
    assert sorted(sd.keys()) == ref_keys
    for k in ref_keys:
        assert tuple(sd[k].shape) == tuple(g['sd.' + k].shape), k
    m.load_state_dict({k: g['sd.' + k].float() for k in ref_keys}, strict=True)


def test_init_matches_reference_reset_parameters():
    g = load_golden('module_init_bias')
    for key, bias in g.items():
        d, L, M, P = [int(x) for x in key.split('_')[1:]]
        m = MSDeformAttn(d, L, M, P, 1.0)
        torch.testing.assert_close(m.sampling_offsets.bias.detach(), bias, rtol=0, atol=0)
        assert m.sampling_offsets.weight.abs().max() == 0
        assert m.attention_weights.weight.abs().max() == 0 and m.attention_weights.bias.abs().max() == 0
        assert m.value_proj.bias.abs().max() == 0 and m.output_proj.bias.abs().max() == 0
        assert m.im2col_step == 64


def test_constructor_errors_and_warning():
    with pytest.raises(ValueError, match='divisible'):
        MSDeformAttn(d_model=100, n_heads=8)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        MSDeformAttn(d_model=96, n_heads=4)  # 24 per head: not a power of two
        assert any('power of 2' in str(x.message) for x in w)
    m = MSDeformAttn(d_model=1024, n_levels=3, n_heads=16, n_points=4, ratio=0.5)
    assert m.value_proj.out_features == 512 and m.output_proj.in_features == 512
    assert m.sampling_offsets.out_features == 16 * 3 * 4 * 2 and m.attention_weights.out_features == 16 * 3 * 4


def test_bad_reference_point_dim_raises():
    m = MSDeformAttn(32, 1, 2, 2)
    shapes = torch.as_tensor([(2, 2)], dtype=torch.long)
    with pytest.raises(ValueError, match='2 or 4'):
        m(torch.zeros(1, 3, 32), torch.zeros(1, 3, 1, 3), torch.zeros(1, 4, 32), shapes, torch.zeros(1, dtype=torch.long))


def test_len_in_assert():
    m = MSDeformAttn(32, 1, 2, 2)
    shapes = torch.as_tensor([(2, 2)], dtype=torch.long)
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 3, 32), torch.zeros(1, 3, 1, 2), torch.zeros(1, 5, 32), shapes, torch.zeros(1, dtype=torch.long))


def test_tensor_memo_and_module_pickle_after_use():
    """the shape-check memo holds weak references with callbacks; it must pickle (as an empty memo) so that
    torch.save(model) / copy.deepcopy / mp.spawn work after the first forward, as they do for the reference module."""
    import copy
    import io
    import pickle
    from vit_adapter_b200 import _cabi
    from vit_adapter_b200.modules import MSDeformAttn
    m = MSDeformAttn(d_model=64, n_levels=2, n_heads=4, n_points=2)
    shapes = torch.as_tensor([(4, 4), (2, 2)], dtype=torch.long)
    m._check_len_in(shapes, 20)                      # what the first forward does: populates the memo
    assert m._checked_shapes.get(shapes, 20)
    m2 = pickle.loads(pickle.dumps(m))
    assert isinstance(m2._checked_shapes, _cabi.TensorMemo) and m2._checked_shapes.get(shapes, 20) is None
    m3 = copy.deepcopy(m)
    assert m3._checked_shapes.get(shapes, 20) is None and m3._merged_cache is None
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m4 = torch.load(buf, weights_only=False)
    assert torch.equal(m4.sampling_offsets.bias, m.sampling_offsets.bias)
    m4._check_len_in(shapes, 20)                     # and the restored memo works
    assert m4._checked_shapes.get(shapes, 20)
