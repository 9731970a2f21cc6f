"""SyncBatchNormNoHostSync (adapter/sync_batchnorm.py) against torch's nn.SyncBatchNorm on 2 GPUs: bit-identical outputs,
gradients and running statistics, and - the point of the module - no device->host synchronisation in the forward."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from torch import nn
    from vit_adapter_b200.adapter import SyncBatchNormNoHostSync
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        torch.manual_seed(7)
        ref = nn.SyncBatchNorm(24).cuda()
        ours = SyncBatchNormNoHostSync(24).cuda()
        with torch.no_grad():
            ref.weight.uniform_(0.5, 1.5); ref.bias.uniform_(-0.5, 0.5)
        ours.load_state_dict(ref.state_dict(), strict=True)
        torch.manual_seed(100 + rank)                              # different data on every rank, different batch sizes
        x = torch.randn(2 + rank, 24, 9, 7, device='cuda')
        gy = torch.randn_like(x)
        xa = x.clone().requires_grad_()
        xb = x.clone().requires_grad_()
        ya = ref(xa)
        ya.backward(gy)
        torch.cuda.synchronize()
        torch.cuda.set_sync_debug_mode('error')                    # any device->host synchronisation raises from here on
        try:
            yb = ours(xb)
        finally:
            torch.cuda.set_sync_debug_mode('default')
        yb.backward(gy)
        torch.cuda.synchronize()
        ok = (torch.equal(ya, yb) and torch.equal(xa.grad, xb.grad) and torch.equal(ref.weight.grad, ours.weight.grad)
              and torch.equal(ref.bias.grad, ours.bias.grad) and torch.equal(ref.running_mean, ours.running_mean)
              and torch.equal(ref.running_var, ours.running_var) and int(ours.num_batches_tracked) == 1)
        # eval mode and single-process use fall back to the parent class
        ours.eval()
        ok = ok and torch.equal(ours(x), ref.eval()(x))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_sync_batchnorm_no_host_sync_matches_torch():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29531, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, True), (1, True)], got


def test_sync_batchnorm_no_host_sync_single_process_is_batchnorm():
    from torch import nn
    from vit_adapter_b200.adapter import SyncBatchNormNoHostSync
    m = SyncBatchNormNoHostSync(8).cuda()
    ref = nn.BatchNorm2d(8).cuda()
    x = torch.randn(4, 8, 5, 5, device='cuda')
    torch.testing.assert_close(m(x), ref(x))
    assert set(m.state_dict()) == set(nn.SyncBatchNorm(8).state_dict())
