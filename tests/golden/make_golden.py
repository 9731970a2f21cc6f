"""Generate the golden fixtures in tests/golden/ by running the REAL reference code.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py

It imports, unmodified and from where they lie,
  /root/reference/detection/ops/functions/ms_deform_attn_func.py   (ms_deform_attn_core_pytorch)
  /root/reference/detection/ops/modules/ms_deform_attn.py          (MSDeformAttn)
  /root/reference/detection/mmdet_custom/models/backbones/adapter_modules.py
        (deform_inputs, Injector, Extractor, InteractionBlock)
with two sys.modules stubs for packages absent here: the compiled
`MultiScaleDeformableAttention` extension and `timm.models.layers.DropPath`; the reference's
MSDeformAttnFunction is routed to its own pure-PyTorch core (the function the reference itself calls
its debug/test oracle). Inputs AND outputs are stored, so the tests do not depend on torch's RNG.
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def load_reference():
    warnings.simplefilter('ignore')
    sys.modules['MultiScaleDeformableAttention'] = types.ModuleType('MultiScaleDeformableAttention')
    timm = types.ModuleType('timm')
    timm_models = types.ModuleType('timm.models')
    timm_layers = types.ModuleType('timm.models.layers')

    class DropPath(torch.nn.Module):  # identity at drop_prob = 0 / eval, which is all the fixtures use
        def __init__(self, drop_prob=0.):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            assert self.drop_prob == 0. or not self.training
            return x

    timm_layers.DropPath = DropPath
    timm.models = timm_models
    timm_models.layers = timm_layers
    sys.modules.update({'timm': timm, 'timm.models': timm_models, 'timm.models.layers': timm_layers})
    sys.path.insert(0, os.path.join(REF, 'detection'))
    import ops.functions.ms_deform_attn_func as ref_func
    import ops.modules.ms_deform_attn as ref_mod

    class _CoreFunction:
        @staticmethod
        def apply(value, shapes, lsi, loc, aw, step):
            return ref_func.ms_deform_attn_core_pytorch(value, shapes, loc, aw)

    ref_mod.MSDeformAttnFunction = _CoreFunction
    spec = importlib.util.spec_from_file_location(
        'ref_adapter_modules',
        os.path.join(REF, 'detection/mmdet_custom/models/backbones/adapter_modules.py'))
    ref_adapter = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_adapter)
    return ref_func, ref_mod, ref_adapter


def level_start(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def op_case(ref_func, name, N, M, D, Lq, shapes, P, seed, dist):
    """One operator-level fixture: inputs, fp64 + fp32 outputs, fp64 grads (autograd through the core)."""
    torch.manual_seed(seed)
    shapes = torch.as_tensor(shapes, dtype=torch.long)
    L = shapes.shape[0]
    S = int(shapes.prod(1).sum())
    if dist == 'ref_test':  # detection/ops/test.py:28-33
        value = torch.rand(N, S, M, D) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        aw = torch.rand(N, Lq, M, L, P) + 1e-5
        aw /= aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    elif dist == 'edges':   # out-of-range samples, exact texel centres, 0 and 1 (every validity branch)
        value = torch.randn(N, S, M, D)
        loc = torch.rand(N, Lq, M, L, P, 2) * 1.2 - 0.1
        flat = loc.view(-1, 2)
        k = flat.shape[0]
        W0 = float(shapes[0, 1])
        flat[0::7] = (torch.randint(0, int(W0), (len(flat[0::7]), 2)).float() + 0.5) / W0
        flat[1::11] = 0.0
        flat[2::13] = 1.0
        flat[3::17] = torch.tensor([0.5 / W0, 1.0 - 0.5 / W0])
        assert k > 17
        aw = torch.softmax(torch.randn(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
    else:
        raise ValueError(dist)
    grad_out = torch.randn(N, Lq, M * D)
    out32 = ref_func.ms_deform_attn_core_pytorch(value, shapes, loc, aw)
    v = value.double().requires_grad_()
    l = loc.double().requires_grad_()
    a = aw.double().requires_grad_()
    out64 = ref_func.ms_deform_attn_core_pytorch(v, shapes, l, a)
    out64.backward(grad_out.double())
    np.savez_compressed(
        os.path.join(HERE, name + '.npz'),
        value=value.numpy(), shapes=shapes.numpy(), lsi=level_start(shapes).numpy(), loc=loc.numpy(),
        aw=aw.numpy(), grad_out=grad_out.numpy(), out_f32=out32.numpy(), out_f64=out64.detach().numpy(),
        grad_value_f64=v.grad.numpy(), grad_loc_f64=l.grad.numpy(), grad_aw_f64=a.grad.numpy())
    print(name, 'S=%d pts=%d' % (S, N * Lq * M * L * P))


def module_case(ref_mod, name, d_model, n_levels, n_heads, n_points, ratio, N, Lq, shapes, seed, ref_dim=2):
    torch.manual_seed(seed)
    m = ref_mod.MSDeformAttn(d_model, n_levels, n_heads, n_points, ratio).double()
    with torch.no_grad():  # move off the all-zero init so every parameter matters
        for p in m.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    shapes = torch.as_tensor(shapes, dtype=torch.long)
    S = int(shapes.prod(1).sum())
    query = torch.randn(N, Lq, d_model, dtype=torch.double)
    feat = torch.randn(N, S, d_model, dtype=torch.double)
    ref_pts = torch.rand(N, Lq, n_levels, ref_dim, dtype=torch.double)
    if ref_dim == 4:
        ref_pts[..., 2:] = ref_pts[..., 2:] * 0.3 + 0.05
    mask = torch.zeros(N, S, dtype=torch.bool)
    mask[:, -3:] = True
    out = m(query, ref_pts, feat, shapes, level_start(shapes), mask)
    out_nomask = m(query, ref_pts, feat, shapes, level_start(shapes), None)
    arrays = {('sd.' + k): v.numpy() for k, v in m.state_dict().items()}
    arrays.update(query=query.numpy(), feat=feat.numpy(), ref_pts=ref_pts.numpy(), shapes=shapes.numpy(),
                  mask=mask.numpy(), out=out.detach().numpy(), out_nomask=out_nomask.detach().numpy(),
                  cfg=np.array([d_model, n_levels, n_heads, n_points], dtype=np.int64), ratio=np.array(ratio))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **arrays)
    print(name)


def module_grad_case(ref_mod, name, d_model, n_levels, n_heads, n_points, ratio, N, Lq, shapes, seed, ref_levels):
    """The reference MSDeformAttn module forward AND backward (fp64 autograd through its own pure-PyTorch core):
    gradients w.r.t. query, feat and every parameter, for a given grad_out. Reference points as the adapter passes them
    ([1, Lq, 1, 2], broadcast over batch and levels) or per level."""
    torch.manual_seed(seed)
    m = ref_mod.MSDeformAttn(d_model, n_levels, n_heads, n_points, ratio).double()
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    shapes = torch.as_tensor(shapes, dtype=torch.long)
    S = int(shapes.prod(1).sum())
    query = torch.randn(N, Lq, d_model, dtype=torch.double, requires_grad=True)
    feat = torch.randn(N, S, d_model, dtype=torch.double, requires_grad=True)
    ref_pts = torch.rand(1, Lq, ref_levels, 2, dtype=torch.double)
    grad_out = torch.randn(N, Lq, d_model, dtype=torch.double)
    out = m(query, ref_pts, feat, shapes, level_start(shapes), None)
    out.backward(grad_out)
    arrays = {('sd.' + k): v.numpy() for k, v in m.state_dict().items()}
    arrays.update({('grad.' + k): v.grad.numpy() for k, v in m.named_parameters()})
    arrays.update(query=query.detach().numpy(), feat=feat.detach().numpy(), ref_pts=ref_pts.numpy(), shapes=shapes.numpy(),
                  out=out.detach().numpy(), grad_out=grad_out.numpy(), grad_query=query.grad.numpy(), grad_feat=feat.grad.numpy(),
                  cfg=np.array([d_model, n_levels, n_heads, n_points], dtype=np.int64), ratio=np.array(ratio))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **arrays)
    print(name)


def load_reference_seg_adapter():
    """segmentation/mmseg_custom/models/backbones/adapter_modules.py: the copy that has InteractionBlockWithCls."""
    sys.path.insert(0, os.path.join(REF, 'segmentation'))
    spec = importlib.util.spec_from_file_location(
        'ref_seg_adapter_modules', os.path.join(REF, 'segmentation/mmseg_custom/models/backbones/adapter_modules.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def vit_block_stand_in(x, H, W):
    """A deterministic, parameter-free stand-in for the ViT blocks between Injector and Extractor (the trunk is out of
    scope); it mixes the class token into the patch tokens so that the cls re-attachment order matters."""
    return x * 1.25 + 0.5 * x.mean(1, keepdim=True) + 0.01 * (H - W)


def adapter_cls_case(ref_seg_adapter, name, dim, heads, ratio, H, W, N, seed):
    """InteractionBlockWithCls.forward of the segmentation copy (:194-234): injector -> cat(cls, x) -> blocks ->
    split -> extractor + 2 extra extractors."""
    torch.manual_seed(seed)
    blk = ref_seg_adapter.InteractionBlockWithCls(dim=dim, num_heads=heads, n_points=4, init_values=0., deform_ratio=ratio,
                                                  extra_extractor=True, with_cffn=True, cffn_ratio=0.25).double()
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    img = torch.zeros(N, 3, H, W)
    di1, di2 = ref_seg_adapter.deform_inputs(img)
    h, w = H // 16, W // 16
    x = torch.randn(N, h * w, dim, dtype=torch.double)
    c = torch.randn(N, (2 * h) * (2 * w) + h * w + (h // 2) * (w // 2), dim, dtype=torch.double)
    cls = torch.randn(N, 1, dim, dtype=torch.double)
    di1d = [di1[0].double(), di1[1], di1[2]]
    di2d = [di2[0].double(), di2[1], di2[2]]
    xo, co, clso = blk(x, c, cls, [vit_block_stand_in, vit_block_stand_in], di1d, di2d, h, w)
    arrays = {('sd.' + k): v.numpy() for k, v in blk.state_dict().items()}
    arrays.update(x=x.numpy(), c=c.numpy(), cls=cls.numpy(), x_out=xo.detach().numpy(), c_out=co.detach().numpy(),
                  cls_out=clso.detach().numpy(), ref1=di1[0].numpy(), shapes1=di1[1].numpy(), lsi1=di1[2].numpy(),
                  ref2=di2[0].numpy(), shapes2=di2[1].numpy(), lsi2=di2[2].numpy(),
                  cfg=np.array([dim, heads, H, W, N], dtype=np.int64), ratio=np.array(ratio))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **arrays)
    print(name)


def load_reference_wsdm_adapter():
    """wsdm2023/mmdet_custom/models/backbones/adapter_modules.py: the copy whose InteractionBlock threads text tokens."""
    sys.path.insert(0, os.path.join(REF, 'wsdm2023'))
    spec = importlib.util.spec_from_file_location(
        'ref_wsdm_adapter_modules', os.path.join(REF, 'wsdm2023/mmdet_custom/models/backbones/adapter_modules.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def text_block_stand_in(x, q, q_mask, H, W):
    """Parameter-free stand-in for the wsdm2023 ViT blocks, which take and return (image tokens, text tokens): the two
    streams exchange their masked means so that the order of the calls and the returned q both matter."""
    m = q_mask.to(q.dtype).unsqueeze(-1)
    qm = (q * m).sum(1, keepdim=True) / m.sum(1, keepdim=True).clamp_min(1)
    return x * 1.125 + 0.25 * qm + 0.01 * (H - W), q * 0.75 + 0.5 * x.mean(1, keepdim=True)


def adapter_text_case(ref_wsdm_adapter, name, dim, heads, ratio, H, W, N, T, seed):
    """InteractionBlock.forward of the wsdm2023 copy (:183-198): injector -> blocks(x, q, q_mask) -> extractors."""
    torch.manual_seed(seed)
    blk = ref_wsdm_adapter.InteractionBlock(dim=dim, num_heads=heads, n_points=4, init_values=0., deform_ratio=ratio,
                                            extra_extractor=True, with_cffn=True, cffn_ratio=0.25).double()
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    img = torch.zeros(N, 3, H, W)
    di1, di2 = ref_wsdm_adapter.deform_inputs(img)
    h, w = H // 16, W // 16
    x = torch.randn(N, h * w, dim, dtype=torch.double)
    c = torch.randn(N, (2 * h) * (2 * w) + h * w + (h // 2) * (w // 2), dim, dtype=torch.double)
    q = torch.randn(N, T, dim, dtype=torch.double)
    q_mask = torch.ones(N, T, dtype=torch.bool)
    q_mask[0, T // 2:] = False
    di1d = [di1[0].double(), di1[1], di1[2]]
    di2d = [di2[0].double(), di2[1], di2[2]]
    xo, co, qo = blk(x, c, q, q_mask, [text_block_stand_in, text_block_stand_in], di1d, di2d, h, w)
    arrays = {('sd.' + k): v.numpy() for k, v in blk.state_dict().items()}
    arrays.update(x=x.numpy(), c=c.numpy(), q=q.numpy(), q_mask=q_mask.numpy(), x_out=xo.detach().numpy(),
                  c_out=co.detach().numpy(), q_out=qo.detach().numpy(), ref1=di1[0].numpy(), shapes1=di1[1].numpy(),
                  lsi1=di1[2].numpy(), ref2=di2[0].numpy(), shapes2=di2[1].numpy(), lsi2=di2[2].numpy(),
                  cfg=np.array([dim, heads, H, W, N], dtype=np.int64), ratio=np.array(ratio))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **arrays)
    print(name)


def init_case(ref_mod, name):
    """The reference's deterministic init of sampling_offsets.bias for the adapter head counts."""
    arrays = {}
    for (d, L, M, P) in [(384, 3, 6, 4), (768, 1, 12, 4), (1024, 3, 16, 4), (256, 4, 8, 4)]:
        m = ref_mod.MSDeformAttn(d, L, M, P, 1.0)
        arrays['bias_%d_%d_%d_%d' % (d, L, M, P)] = m.sampling_offsets.bias.detach().numpy()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **arrays)
    print(name)


def adapter_case(ref_adapter, name, dim, heads, ratio, H, W, N, seed):
    """deform_inputs + one InteractionBlock (Injector -> [no ViT blocks] -> Extractor + 2 extra extractors)."""
    torch.manual_seed(seed)
    blk = ref_adapter.InteractionBlock(dim=dim, num_heads=heads, n_points=4, init_values=0.,
                                       deform_ratio=ratio, extra_extractor=True, with_cffn=True,
                                       cffn_ratio=0.25).double()
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    img = torch.zeros(N, 3, H, W)
    di1, di2 = ref_adapter.deform_inputs(img)
    h, w = H // 16, W // 16
    x = torch.randn(N, h * w, dim, dtype=torch.double)
    c = torch.randn(N, (2 * h) * (2 * w) + h * w + (h // 2) * (w // 2), dim, dtype=torch.double)
    di1d = [di1[0].double(), di1[1], di1[2]]
    di2d = [di2[0].double(), di2[1], di2[2]]
    xo, co = blk(x, c, [], di1d, di2d, h, w)
    arrays = {('sd.' + k): v.numpy() for k, v in blk.state_dict().items()}
    arrays.update(x=x.numpy(), c=c.numpy(), x_out=xo.detach().numpy(), c_out=co.detach().numpy(),
                  ref1=di1[0].numpy(), shapes1=di1[1].numpy(), lsi1=di1[2].numpy(),
                  ref2=di2[0].numpy(), shapes2=di2[1].numpy(), lsi2=di2[2].numpy(),
                  cfg=np.array([dim, heads, H, W, N], dtype=np.int64), ratio=np.array(ratio))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **arrays)
    print(name)


def dwconv_case(ref_adapter, name, C, H, W, B, seed):
    """The reference's DWConv module on a packed [B, 21n, C] token sequence: output and autograd gradients (fp64)."""
    torch.manual_seed(seed)
    m = ref_adapter.DWConv(C).double()
    n = (H // 2) * (W // 2)
    x = torch.randn(B, 21 * n, C, dtype=torch.double, requires_grad=True)
    gy = torch.randn(B, 21 * n, C, dtype=torch.double)
    y = m(x, H, W)
    y.backward(gy)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), x=x.detach().numpy(), weight=m.dwconv.weight.detach().numpy(),
                        bias=m.dwconv.bias.detach().numpy(), y=y.detach().numpy(), grad_y=gy.numpy(), grad_x=x.grad.numpy(),
                        grad_weight=m.dwconv.weight.grad.numpy(), grad_bias=m.dwconv.bias.grad.numpy(),
                        cfg=np.array([C, H, W, B], dtype=np.int64))
    print(name)


def layernorm_case(ref_adapter, name, C, shape, seed, heads=4):
    """The `query_norm` LayerNorm that the reference's Injector constructs (adapter_modules.py:127-134: norm_layer =
    partial(nn.LayerNorm, eps=1e-6)), with non-trivial affine parameters: output and autograd gradients (fp64)."""
    torch.manual_seed(seed)
    inj = ref_adapter.Injector(dim=C, num_heads=heads, n_points=4, n_levels=3).double()
    norm = inj.query_norm
    with torch.no_grad():
        norm.weight.copy_(1.0 + 0.3 * torch.randn(C, dtype=torch.double))
        norm.bias.copy_(0.2 * torch.randn(C, dtype=torch.double))
    x = (2.0 * torch.randn(*shape, C, dtype=torch.double) + 0.7).requires_grad_()
    gy = torch.randn(*shape, C, dtype=torch.double)
    y = norm(x)
    y.backward(gy)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), x=x.detach().numpy(), weight=norm.weight.detach().numpy(),
                        bias=norm.bias.detach().numpy(), y=y.detach().numpy(), grad_y=gy.numpy(), grad_x=x.grad.numpy(),
                        grad_weight=norm.weight.grad.numpy(), grad_bias=norm.bias.grad.numpy(),
                        eps=np.array([norm.eps], dtype=np.float64))
    print(name)


def main():
    ref_func, ref_mod, ref_adapter = load_reference()
    if 'round2' in sys.argv[1:]:      # only the fixtures added in round 2 (the others are unchanged)
        module_grad_case(ref_mod, 'module_l3_grads', d_model=128, n_levels=3, n_heads=4, n_points=4, ratio=1.0, N=2, Lq=16,
                         shapes=[(8, 8), (4, 4), (2, 2)], seed=61, ref_levels=1)
        module_grad_case(ref_mod, 'module_l1_grads', d_model=128, n_levels=1, n_heads=4, n_points=4, ratio=1.0, N=2, Lq=84,
                         shapes=[(4, 4)], seed=62, ref_levels=1)
        adapter_cls_case(load_reference_seg_adapter(), 'adapter_block_cls', dim=32, heads=4, ratio=0.5, H=64, W=96, N=2, seed=63)
        return
    if 'text' in sys.argv[1:]:        # only the wsdm2023 text-token interaction block (added late in round 2)
        adapter_text_case(load_reference_wsdm_adapter(), 'adapter_block_text', dim=32, heads=4, ratio=0.5, H=64, W=96, N=2, T=5, seed=64)
        return
    if 'layernorm' in sys.argv[1:]:   # only the N2 fixtures (the others are unchanged)
        layernorm_case(ref_adapter, 'layernorm_c96', C=96, shape=(2, 21), seed=51)
        layernorm_case(ref_adapter, 'layernorm_c768', C=768, shape=(1, 5), seed=52, heads=12)
        return
    # 1. the reference's own test fixture (detection/ops/test.py:16-37), seed 3 == SURVEY App. A.4 KAT
    op_case(ref_func, 'op_kat_seed3', N=1, M=2, D=2, Lq=2, shapes=[(6, 4), (3, 2)], P=2, seed=3, dist='ref_test')
    # 2. injector-like (L=3) and extractor-like (L=1) miniatures, with the edge distribution
    op_case(ref_func, 'op_inj_edges', N=2, M=3, D=8, Lq=16, shapes=[(8, 8), (4, 4), (2, 2)], P=4, seed=11, dist='edges')
    op_case(ref_func, 'op_ext_edges', N=2, M=3, D=8, Lq=84, shapes=[(4, 4)], P=4, seed=12, dist='edges')
    # 3. non-square levels, odd channel count (generic kernel), ref-test distribution
    op_case(ref_func, 'op_odd_d5', N=1, M=2, D=5, Lq=7, shapes=[(5, 3), (2, 7)], P=3, seed=13, dist='ref_test')
    # 4. vector-kernel channel counts with the real head widths
    op_case(ref_func, 'op_d32_edges', N=1, M=2, D=32, Lq=9, shapes=[(6, 6), (3, 3), (2, 2)], P=4, seed=14, dist='edges')
    op_case(ref_func, 'op_d64_edges', N=1, M=2, D=64, Lq=9, shapes=[(5, 7)], P=4, seed=15, dist='edges')
    # 5. module level
    module_case(ref_mod, 'module_l3', d_model=32, n_levels=3, n_heads=4, n_points=4, ratio=1.0, N=2, Lq=10,
                shapes=[(6, 6), (3, 3), (2, 2)], seed=21)
    module_case(ref_mod, 'module_ratio_half_box', d_model=32, n_levels=2, n_heads=2, n_points=2, ratio=0.5, N=1,
                Lq=6, shapes=[(4, 5), (2, 3)], seed=22, ref_dim=4)
    init_case(ref_mod, 'module_init_bias')
    # 6. adapter level
    adapter_case(ref_adapter, 'adapter_block', dim=32, heads=4, ratio=0.5, H=64, W=96, N=2, seed=31)
    # 7. ConvFFN's depth-wise conv on the token layout (C = 12: vector path for fp32; C = 6: scalar path)
    dwconv_case(ref_adapter, 'dwconv_tokens', C=12, H=4, W=6, B=2, seed=41)
    dwconv_case(ref_adapter, 'dwconv_tokens_c6', C=6, H=2, W=2, B=1, seed=42)
    # 8. the LayerNorm in front of the adapter's Linears
    layernorm_case(ref_adapter, 'layernorm_c96', C=96, shape=(2, 21), seed=51)
    layernorm_case(ref_adapter, 'layernorm_c768', C=768, shape=(1, 5), seed=52, heads=12)
    # 9. round 2: module gradients at head widths the fused kernels take (D = 32, three levels and one level), and
    #    the class-token interaction block of the segmentation copy
    module_grad_case(ref_mod, 'module_l3_grads', d_model=128, n_levels=3, n_heads=4, n_points=4, ratio=1.0, N=2, Lq=16,
                     shapes=[(8, 8), (4, 4), (2, 2)], seed=61, ref_levels=1)
    module_grad_case(ref_mod, 'module_l1_grads', d_model=128, n_levels=1, n_heads=4, n_points=4, ratio=1.0, N=2, Lq=84,
                     shapes=[(4, 4)], seed=62, ref_levels=1)
    adapter_cls_case(load_reference_seg_adapter(), 'adapter_block_cls', dim=32, heads=4, ratio=0.5, H=64, W=96, N=2, seed=63)
    # 10. the wsdm2023 interaction block (text tokens threaded through the ViT blocks)
    adapter_text_case(load_reference_wsdm_adapter(), 'adapter_block_text', dim=32, heads=4, ratio=0.5, H=64, W=96, N=2, T=5, seed=64)


if __name__ == '__main__':
    main()
