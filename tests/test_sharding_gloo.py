"""N>1 path on CPU: world_size-2 gloo. Checks the batch sharding used by bench.py --gpus N: shards
cover the batch exactly once, the op is batch-separable (sharded results concatenate to the unsharded
result, verified with the C oracle as the per-rank compute stand-in), and the timing reduction is a MAX."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, make_inputs


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import c_oracle
    from vit_adapter_b200.sharding import batch_shard, max_over_ranks, shard_op_inputs
    inp = make_inputs(5, 3, 8, 12, [(6, 6), (3, 3)], 4, seed=4, dist='edges')  # identical on every rank (seeded)
    b, e = batch_shard(5, world, rank)
    v, l, a, go = shard_op_inputs(inp['value'], inp['loc'], inp['aw'], world, rank, inp['grad_out'])
    assert v.shape[0] == e - b
    out = c_oracle.forward(v, inp['shapes'], inp['lsi'], l, a)
    gv, gl, ga = c_oracle.backward(v, inp['shapes'], inp['lsi'], l, a, go)
    sizes = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([e - b]))
    assert sum(int(s) for s in sizes) == 5
    # gather padded shards on rank 0 and compare with the unsharded run
    pad = torch.zeros((3,) + out.shape[1:])
    pad[:e - b] = out
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    padg = torch.zeros((3,) + gv.shape[1:])
    padg[:e - b] = gv
    gvs = [torch.zeros_like(padg) for _ in range(world)]
    dist.all_gather(gvs, padg)
    if rank == 0:
        full = c_oracle.forward(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'])
        fgv, _, _ = c_oracle.backward(inp['value'], inp['shapes'], inp['lsi'], inp['loc'], inp['aw'], inp['grad_out'])
        cat = torch.cat([o[:int(s)] for o, s in zip(outs, sizes)], 0)
        catg = torch.cat([o[:int(s)] for o, s in zip(gvs, sizes)], 0)
        assert torch.equal(cat, full) and torch.equal(catg, fgv)
    m = max_over_ranks(10.0 + rank)
    assert m == 10.0 + world - 1
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, 'ok%d' % rank), 'w').write('ok')


def test_batch_shard_partition():
    from vit_adapter_b200.sharding import batch_shard
    for total in (1, 2, 5, 16, 17):
        for world in (1, 2, 3, 8):
            spans = [batch_shard(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        batch_shard(4, 2, 2)


@pytest.mark.timeout(120)
def test_two_rank_gloo(tmp_path):
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / 'ok0').exists() and (tmp_path / 'ok1').exists()
