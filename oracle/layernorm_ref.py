"""CPU restatement of the LayerNorm in front of the adapter's Linears. TEST INFRASTRUCTURE ONLY.

Follows what the reference calls at adapter_modules.py:110-116,142-145 (`self.query_norm(query)`, `self.feat_norm(feat)`,
`self.ffn_norm(query)` with norm_layer = partial(nn.LayerNorm, eps=1e-6), :93,:130): per row of C channels
    y = (x - mean) / sqrt(var + eps) * weight + bias          (biased variance)
and its analytic gradients. Pinned by tests/test_layernorm.py against golden vectors produced by the LayerNorm module that
the reference's own Injector constructs (tests/golden/make_golden.py::layernorm_case).
"""
import torch


def layernorm(x, weight, bias, eps):
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    y = (x - mean) / torch.sqrt(var + eps) * weight
    return y + bias if bias is not None else y


def layernorm_backward(x, weight, eps, grad_y):
    """(grad_x, grad_weight, grad_bias)."""
    C = x.shape[-1]
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    rstd = 1.0 / torch.sqrt(var + eps)
    xhat = (x - mean) * rstd
    g = grad_y * weight
    gx = rstd * (g - g.mean(-1, keepdim=True) - xhat * (g * xhat).mean(-1, keepdim=True))
    flat_gy, flat_xh = grad_y.reshape(-1, C), xhat.reshape(-1, C)
    return gx, (flat_gy * flat_xh).sum(0), flat_gy.sum(0)
