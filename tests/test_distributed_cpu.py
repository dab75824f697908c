"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sequence sharding, max-over-ranks timing,
and the single flat gradient all-reduce of the training path."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from aum_b200 import dist as D
    assert D.init_from_env("gloo") == world
    # timing rule: max over ranks
    assert D.max_over_ranks(10.0 + rank, "cpu") == 10.0 + world - 1
    # sequence sharding covers every item exactly once
    lo, hi = D.shard_range(129, rank, world)
    cover = torch.zeros(129)
    cover[lo:hi] = 1
    dist.all_reduce(cover)
    assert (cover == 1).all()
    # flat gradient all-reduce == mean of per-rank gradients, gradients alias the flat buffer
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    red = D.FlatGradReducer(lin.parameters())
    x = torch.full((4, 5), float(rank + 1))
    lin(x).sum().backward()
    local = [p.grad.clone() for p in lin.parameters()]
    assert all(p.grad.data_ptr() >= red.flat.data_ptr() for p in lin.parameters())
    red.reduce()
    gathered = [torch.zeros_like(torch.cat([g.flatten() for g in local])) for _ in range(world)]
    dist.all_gather(gathered, torch.cat([g.flatten() for g in local]))
    # (every tensor starts on an aligned offset of the flat buffer; the padding in between stays zero)
    mean = torch.stack(gathered).mean(0)
    packed = torch.cat([red.flat[o:o + p.numel()] for p, o in zip(red.params, red.offsets)])
    torch.testing.assert_close(packed, mean)
    assert all(o % D.FlatGradReducer.ALIGN == 0 for o in red.offsets)
    pad = torch.ones_like(red.flat, dtype=torch.bool)
    for p, o in zip(red.params, red.offsets):
        pad[o:o + p.numel()] = False
    assert (red.flat[pad] == 0).all()
    red.zero()
    assert all((p.grad == 0).all() for p in lin.parameters())
    # chunked, overlapped form: parameters in gradient-ready order (last layer first), the first chunk's all-reduce
    # launched from a tensor hook in the middle of backward, the rest by reduce(); same averaged gradients
    torch.manual_seed(0)
    l1, l2 = torch.nn.Linear(5, 7), torch.nn.Linear(7, 3)
    order = list(l2.parameters()) + list(l1.parameters())
    red2 = D.FlatGradReducer(order, chunk_after=[2])
    assert red2.n_chunks == 2 and red2.bounds[1] == red2.offsets[2]
    h = l1(x)
    h.register_hook(red2.hook(0))          # backward passes here once l2's gradients are final
    l2(h).sum().backward()
    assert red2.async_launches == 1        # chunk 0 went out during backward
    red2.reduce()
    assert red2.async_launches == 2
    packed2 = torch.cat([red2.flat[o:o + p.numel()] for p, o in zip(red2.params, red2.offsets)])
    ref_order = torch.cat([mean_part for mean_part in (
        torch.cat([packed[sum(q.numel() for q in list(lin.parameters())[:i]):][:p_.numel()]
                   for i, p_ in enumerate(lin.parameters()) if i in idx]) for idx in ((2, 3), (0, 1)))])
    torch.testing.assert_close(packed2, ref_order)
    dist.destroy_process_group()
    ret[rank] = 1


def test_two_rank_gloo_host_logic():
    world = 2
    port = 29500 + (os.getpid() % 500)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: 1, 1: 1}
