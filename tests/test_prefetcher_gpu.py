"""tools.utils.CudaPrefetcher: batches arrive intact and in order through the two persistent device slots while the
consumer keeps the GPU busy; slots are reused (no per-batch allocation) and re-allocated when the shape changes."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_prefetcher_delivers_batches_in_order_with_slot_reuse():
    from simpleaicv_pytorch_training_examples_b200.tools.utils import CudaPrefetcher
    g = torch.Generator().manual_seed(0)
    host = [{'image': torch.randn(8, 3, 64, 64, generator=g).pin_memory(), 'label': torch.randint(0, 10, (8,), generator=g).pin_memory(),
             'name': f'b{i}'} for i in range(7)]
    host.append({'image': torch.randn(5, 3, 64, 64, generator=g).pin_memory(), 'label': torch.randint(0, 10, (5,), generator=g).pin_memory(),
                 'name': 'ragged'})
    busy = torch.randn(2048, 2048, device='cuda')
    sums, ptrs = [], []
    for i, batch in enumerate(CudaPrefetcher(host)):
        assert batch['name'] == host[i]['name']
        ptrs.append(batch['image'].data_ptr())
        for _ in range(4):                      # asynchronous work that outlives the Python-side use of the batch
            busy = torch.tanh(busy @ busy * 1e-3)
        sums.append((batch['image'].double().sum() + batch['label'].double().sum()).clone())
    torch.cuda.synchronize()
    assert len(sums) == len(host)
    for s, h in zip(sums, host):
        assert abs(float(s) - float(h['image'].double().sum() + h['label'].double().sum())) < 1e-6
    assert len(set(ptrs[:7])) == 2 and ptrs[0] == ptrs[2] == ptrs[4] and ptrs[1] == ptrs[3]      # two alternating slots
