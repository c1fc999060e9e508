"""World-size-2 CPU (gloo) tests of the batch-sharded data-parallel plumbing:
sharding, rank-0 init + broadcast, flat-bucket gradient all-reduce == single-process gradient on the
concatenated batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import impflow_b200
    par = impflow_b200.parallel
    torch.manual_seed(7 + rank)                       # ranks start with DIFFERENT weights
    model = torch.nn.Sequential(torch.nn.Linear(5, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    versions = [p._version for p in model.parameters()]
    par.broadcast_module(model, 0)                    # ... and end up with rank 0's
    if rank != 0:      # the receive must be visible to the version-keyed host caches (effective weights, sigma)
        assert all(p._version > v for p, v in zip(model.parameters(), versions))
    torch.manual_seed(0)
    x = torch.randn(12, 5)                            # same global batch everywhere
    xs = par.shard_batch(x)
    assert xs.shape[0] == 6 and torch.equal(xs, x[rank * 6:(rank + 1) * 6])
    bucket = par.FlatGradBucket(model.parameters())
    bucket.zero()
    model(xs).pow(2).mean().backward()
    bucket.allreduce_mean()
    flat = torch.cat([v.reshape(-1) for v in bucket.views]).clone()    # the bucket pads every tensor to 256 B
    assert float(bucket.flat.sum()) == float(flat.sum()) or abs(float(bucket.flat.sum()) - float(flat.sum())) < 1e-5
    for p in model.parameters():                      # grads are views into the bucket
        assert p.grad.data_ptr() >= bucket.flat.data_ptr()
    w0 = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    loss = par.allreduce_mean_scalar(model(xs).pow(2).mean().detach().clone())
    q.put((rank, flat, w0, float(loss)))
    dist.destroy_process_group()


def test_two_rank_gradients_match_single_process():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, g0, w0, l0), (_, g1, w1, l1) = results
    assert torch.equal(w0, w1)                        # broadcast made the replicas identical
    torch.testing.assert_close(g0, g1)                # all-reduce result identical on both ranks
    # single-process reference on the full batch with rank 0's weights
    torch.manual_seed(7)
    model = torch.nn.Sequential(torch.nn.Linear(5, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    torch.manual_seed(0)
    x = torch.randn(12, 5)
    full = model(x).pow(2).mean()
    full.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    torch.testing.assert_close(g0, ref, rtol=1e-5, atol=1e-6)
    assert abs(l0 - float(full)) < 1e-6 and abs(l1 - float(full)) < 1e-6


def test_shard_batch_covers_ragged_batches():
    import impflow_b200
    x = torch.arange(10).view(10, 1)
    parts = [impflow_b200.parallel.shard_batch(x, r, 4) for r in range(4)]
    assert [p.shape[0] for p in parts] == [3, 3, 3, 1]
    assert torch.equal(torch.cat(parts), x)
    assert impflow_b200.parallel.shard_batch(x[:2], 3, 4).shape[0] == 0     # empty shard


def test_flat_bucket_gather_semantics():
    """FlatGradBucket: gradients arrive outside the bucket (the .grad slots are emptied by zero()), one gather moves
    them into the flat buffer; parameters without a gradient keep zeros — also after zero_grad(set_to_none=True)
    instead of zero() — and are reported in had_grad; a second gather of the same step is a no-op; gradients whose
    size-1 dimensions carry unusual strides are gathered correctly."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import impflow_b200
    par = impflow_b200.parallel
    torch.manual_seed(1)
    ps = [torch.nn.Parameter(torch.randn(4, 3, 1, 1)), torch.nn.Parameter(torch.randn(5)),
          torch.nn.Parameter(torch.randn(())), torch.nn.Parameter(torch.randn(2, 2))]
    bucket = par.FlatGradBucket(ps)
    assert bucket.flat.numel() % 1024 == 0 and all(o % bucket.ALIGN == 0 for o in bucket.offsets)
    bucket.zero()
    assert all(p.grad is None for p in ps)
    g0 = torch.randn(4, 3).as_strided((4, 3, 1, 1), (3, 1, 7, 7))        # odd strides on the size-1 dimensions
    ps[0].grad, ps[1].grad, ps[3].grad = g0, torch.randn(5), torch.randn(2, 2)
    want = [g0.clone(), ps[1].grad.clone(), torch.zeros(()), ps[3].grad.clone()]
    bucket.gather_strays()
    assert bucket.had_grad == [True, True, False, True]
    for p, v, w in zip(ps, bucket.views, want):
        assert p.grad is v and torch.equal(v, w)
    snapshot = bucket.flat.clone()
    bucket.gather_strays()                                               # e.g. the optimiser after allreduce_mean
    assert torch.equal(bucket.flat, snapshot) and bucket.had_grad == [True, True, False, True]
    # a caller that uses zero_grad(set_to_none=True) instead of zero(): stale values must not survive
    for p in ps:
        p.grad = None
    ps[1].grad = torch.ones(5)
    bucket.gather_strays()
    assert bucket.had_grad == [False, True, False, False]
    assert torch.equal(bucket.views[1], torch.ones(5))
    assert float(bucket.flat.abs().sum()) == 5.0
