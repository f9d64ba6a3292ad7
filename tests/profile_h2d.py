"""Pinned host -> device copy bandwidth of this box: one stream, and the same bytes split over 2 / 4 streams; with and
without a concurrently running GEMM-heavy kernel stream (is the copy slowed down by compute?)."""
import torch

x = torch.randn(256, 3, 224, 224).pin_memory()
d = torch.empty_like(x, device='cuda')
nb = x.numel() * 4
a = torch.randn(8192, 8192, device='cuda', dtype=torch.bfloat16)


def run(parts, busy):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    step = (x.shape[0] + parts - 1) // parts
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for st in streams:
        st.wait_stream(torch.cuda.current_stream())
    for it in range(5):
        if busy:
            for _ in range(4):
                a @ a
        for j, st in enumerate(streams):
            with torch.cuda.stream(st):
                d[j * step:(j + 1) * step].copy_(x[j * step:(j + 1) * step], non_blocking=True)
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / 5


for busy in (False, True):
    for parts in (1, 2, 4):
        ms = run(parts, busy)
        print(f'H2D {nb / 1e6:.0f} MB in {parts} stream(s), compute {"busy" if busy else "idle"}: {ms:.2f} ms per copy round ({nb / ms / 1e6:.1f} GB/s if copy-bound)')
