"""Data-parallel wrapper replacing torch DDP for models executed by the B200 runtime
(reference: tools/utils.py:175-202 build_training_mode wraps the model in
nn.parallel.DistributedDataParallel; tools/scripts.py:168-181 uses ``model.no_sync()``).

One process per GPU.  Parameter gradients are produced by the runtime directly into flat fp32
buckets (reverse registration order, like DDP's reducer); as soon as every gradient of a bucket
has been written, the bucket is all-reduced (average) over NCCL on a side stream so that the
collective overlaps the remaining backward kernels.  There is no per-forward buffer broadcast
and no barrier; replicas start identical because rank 0's parameters and buffers are broadcast
once at construction.  Parameters that receive no gradient in a step are reduced as zeros
(fixed bucket layout).
"""
import contextlib

import torch
import torch.distributed as dist
import torch.nn as nn


class _Bucket:
    def __init__(self, flat, params):
        self.flat = flat
        self.params = params      # [(param, view)]
        self.ready = set()
        self.work = None


class B200DataParallel(nn.Module):

    def __init__(self, module, process_group=None, bucket_cap_mb=25, broadcast_from_rank0=True):
        super().__init__()
        self.module = module
        self.process_group = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.require_sync = True
        params = [p for p in module.parameters() if p.requires_grad]
        assert params, 'no trainable parameters'
        self.device = params[0].device
        if broadcast_from_rank0 and self.world_size > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=dist.get_global_rank(process_group, 0) if process_group else 0,
                               group=process_group)
        sink = module.grad_sink()
        cap = int(bucket_cap_mb * 1024 * 1024) // 4
        self.buckets, self.bucket_of = [], {}
        cur, cur_n = [], 0
        for p in reversed(params):
            if cur and cur_n + p.numel() > cap:
                self._make_bucket(cur, sink)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self._make_bucket(cur, sink)
        sink.on_ready = self._param_ready
        sink.on_backward_end = self._finish
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.device.type == 'cuda' else None

    def _make_bucket(self, params, sink):
        total = sum((p.numel() + 63) // 64 * 64 for p in params)  # 256 B aligned views
        flat = torch.zeros(total, device=self.device, dtype=torch.float32)
        off, pv = 0, []
        for p in params:
            v = flat[off:off + p.numel()].view_as(p)
            sink.views[id(p)] = v
            pv.append((p, v))
            off += (p.numel() + 63) // 64 * 64
        b = _Bucket(flat, pv)
        for p in params:
            self.bucket_of[id(p)] = b
        self.buckets.append(b)

    # ---- hooks fired by the runtime during backward
    def _param_ready(self, p):
        b = self.bucket_of.get(id(p))
        if b is None:
            return
        b.ready.add(id(p))
        if self.require_sync and self.world_size > 1 and len(b.ready) == len(b.params) and b.work is None:
            self._launch(b)

    def _launch(self, b):
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.comm_stream):
                b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.AVG, group=self.process_group, async_op=True)
        else:  # gloo (CPU tests): no AVG
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.process_group, async_op=True)

    def _finish(self):
        if self.require_sync and self.world_size > 1:
            for b in self.buckets:
                if b.work is None:
                    # parameters without a gradient this step take part as zeros
                    for p, v in b.params:
                        if id(p) not in b.ready:
                            v.zero_()
                            if p.grad is None:
                                p.grad = v
                    self._launch(b)
            for b in self.buckets:
                b.work.wait()
                if self.comm_stream is None:
                    b.flat.div_(self.world_size)
                b.work = None
            if self.comm_stream is not None:
                torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
        for b in self.buckets:
            b.ready.clear()

    def reduce_now(self):
        """All-reduce (average) the current contents of every bucket, e.g. after backward passes run
        under no_sync().  Parameters without a gradient take part as zeros."""
        old, self.require_sync = self.require_sync, True
        try:
            for b in self.buckets:
                b.ready.update(id(p) for p, _ in b.params if p.grad is not None)
            self._finish()
        finally:
            self.require_sync = old

    @contextlib.contextmanager
    def no_sync(self):
        """Gradient accumulation: skip the all-reduce for backward passes run inside."""
        old, self.require_sync = self.require_sync, False
        try:
            yield
        finally:
            self.require_sync = old

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)
