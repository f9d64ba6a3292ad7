"""Run under torchrun with >= 2 GPUs: checks distributed.B200DataParallel over NCCL.

Every rank trains resnet18cifar one step on its own batch; the all-reduced gradients must equal
the mean over ranks of the gradients each rank computes alone on the same batches (recomputed
locally, since every rank can regenerate every rank's batch), and replicas must stay identical
after the optimizer step.  Also exercises no_sync() accumulation.
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/ddp_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def batch_of(rank, n=32):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(n, 3, 32, 32, generator=g), torch.randint(0, 100, (n,), generator=g)


def main():
    from simpleaicv_pytorch_training_examples_b200.classification import backbones, losses
    from simpleaicv_pytorch_training_examples_b200.distributed import B200DataParallel
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    crit = losses.CELoss()
    torch.manual_seed(1234 + rank)  # different init per rank: the wrapper must broadcast rank 0's
    model = backbones.resnet18cifar(num_classes=100).cuda().train()
    ddp = B200DataParallel(model, bucket_cap_mb=4)
    assert len(ddp.buckets) > 2
    flat0 = torch.cat([p.detach().flatten() for p in model.parameters()])
    ref0 = flat0.clone()
    dist.broadcast(ref0, src=0)
    assert torch.equal(flat0, ref0), f'rank {rank}: parameters were not broadcast from rank 0'

    # (1) exact check of the bucket / view / all-reduce mechanics: local gradients (no_sync), gathered
    #     and averaged by hand, must equal what reduce_now() leaves in param.grad
    x, y = batch_of(rank)
    with ddp.no_sync():
        crit(ddp(x.cuda()), y.cuda()).backward()
    torch.cuda.synchronize()
    local = torch.cat([p.grad.detach().flatten() for p in model.parameters()])
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = torch.stack(gathered).mean(0)
    ddp.reduce_now()
    torch.cuda.synchronize()
    got = torch.cat([p.grad.detach().flatten() for p in model.parameters()])
    err = ((got - expect).norm() / expect.norm()).item()
    assert err < 1e-6, f'rank {rank}: all-reduced gradients differ from the mean of the local ones: {err}'
    for p in model.parameters():
        p.grad = None

    # (2) overlapped path: a normal backward all-reduces bucket by bucket; afterwards every rank holds
    #     identical, finite gradients and replicas stay identical after the optimizer step
    crit(ddp(x.cuda()), y.cuda()).backward()
    torch.cuda.synchronize()
    g = torch.cat([p.grad.detach().flatten() for p in model.parameters()])
    g0 = g.clone()
    dist.broadcast(g0, src=0)
    assert torch.isfinite(g).all() and torch.equal(g, g0), f'rank {rank}: gradients differ across ranks after backward'
    # the kernels are deterministic (fixed-order reductions, no float atomics) and NCCL reduces a buffer
    # of a given size in a fixed order: the overlapped result must equal the non-overlapped one BIT FOR BIT
    # (a race between the comm stream and the backward kernels would show up here) ...
    from simpleaicv_pytorch_training_examples_b200.distributed import overlap_self_check
    ndiff = overlap_self_check(ddp, lambda: crit(ddp(x.cuda()), y.cuda()).backward())
    assert ndiff == 0, f'rank {rank}: overlapped all-reduce differs from no_sync + reduce_now in {ndiff} elements'
    # ... and equals the hand-averaged gradients of (1) up to NCCL's fp32 averaging order
    rel = ((g - expect).norm() / expect.norm()).item()
    assert rel < 1e-6, f'rank {rank}: overlapped all-reduce differs from the hand-averaged gradients: {rel}'
    crit(ddp(x.cuda()), y.cuda()).backward()
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    opt.step()
    opt.zero_grad()
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    ref_flat = flat.clone()
    dist.broadcast(ref_flat, src=0)
    assert torch.equal(flat, ref_flat), f'rank {rank}: parameters diverged from rank 0'
    # (3) gradient accumulation: a no_sync micro-step followed by a synced one
    with ddp.no_sync():
        crit(ddp(x.cuda()), y.cuda()).backward()
    crit(ddp(x.cuda()), y.cuda()).backward()
    torch.cuda.synchronize()
    ga = torch.cat([p.grad.detach().flatten() for p in model.parameters()])
    ga0 = ga.clone()
    dist.broadcast(ga0, src=0)
    assert torch.isfinite(ga).all() and torch.equal(ga, ga0)
    print(f'rank {rank}: ddp ok, exact all-reduce rel err {err:.2e}, overlapped vs hand-averaged {rel:.2e} (bit-exact vs non-overlapped), '
          f'buckets {len(ddp.buckets)}', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
