"""The adapter interaction (forward + backward + optimizer) records into ONE CUDA graph and replays to the eager result."""
import pytest
import torch

import vit_adapter_b200 as vab


def test_graphed_step_needs_cuda():
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    with pytest.raises(RuntimeError, match='CPU'):
        vab.GraphedStep(lambda: None)


@pytest.mark.gpu
@pytest.mark.parametrize('amp', [False, True], ids=['f32', 'bf16-autocast'])
def test_interaction_block_graph_replay_matches_eager(amp):
    from vit_adapter_b200 import _cabi
    from vit_adapter_b200.adapter import InteractionBlock, deform_inputs
    torch.manual_seed(0)
    dev = torch.device('cuda')
    dim, heads, side, N = 64, 4, 128, 2
    if amp:
        vab.set_amp_value_dtype(torch.bfloat16)
    try:
        h = side // 16

        def build():
            torch.manual_seed(1)
            blk = InteractionBlock(dim, heads, 4, deform_ratio=0.5, cffn_ratio=0.25, init_values=0.3, extra_extractor=True).to(dev)
            with torch.no_grad():
                for p in blk.parameters():
                    p.add_(0.02 * torch.randn_like(p))
            opt = torch.optim.SGD(blk.parameters(), lr=1e-2)
            return blk, opt
        di1, di2 = deform_inputs(torch.zeros(N, 3, side, side, device=dev))
        x = torch.randn(N, h * h, dim, device=dev)
        c = torch.randn(N, 21 * (h // 2) ** 2, dim, device=dev)

        def make_step(blk, opt):
            def step():
                with torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp):
                    xo, co = blk(x, c, [], di1, di2, h, h)
                loss = xo.float().square().mean() + co.float().square().mean()
                opt.zero_grad(set_to_none=False)
                loss.backward()
                opt.step()
                return loss
            return step

        blk_e, opt_e = build()
        eager = make_step(blk_e, opt_e)
        losses_e = [float(eager().detach()) for _ in range(6)]            # 3 warm-up + capture pass + 2 replays below = 6 updates

        blk_g, opt_g = build()
        for p in blk_g.parameters():
            p.grad = torch.zeros_like(p)                         # static gradient buffers for the captured zero_grad
        g = vab.GraphedStep(make_step(blk_g, opt_g), warmup=3)   # capture itself does not execute the kernels
        n0 = _cabi.launch_count()
        l4 = float(g().detach())
        l5 = float(g().detach())
        assert _cabi.launch_count() == n0                        # replays go through the graph, not through the binding
        tol = 5e-2 if amp else 2e-3                              # atomics order + (amp) bf16
        assert abs(l4 - losses_e[3]) <= tol * abs(losses_e[3]), (l4, losses_e)
        assert abs(l5 - losses_e[4]) <= tol * abs(losses_e[4]), (l5, losses_e)
        assert l5 < l4 < losses_e[0]                             # and it is really training
    finally:
        vab.set_amp_value_dtype(torch.float32)
