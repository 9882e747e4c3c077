"""Pinned host -> device copy bandwidth of this box (what bounds the host-buffer path: ~300 KB per window)."""
import torch, time
for mb in (16, 64, 310):
    n = mb * 1024 * 1024 // 8
    h = torch.empty(n, dtype=torch.float64).pin_memory(); d = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"H2D pinned {mb} MB: {ms:.2f} ms  {mb / 1024 / (ms * 1e-3):.1f} GiB/s")
    e0.record()
    for _ in range(5): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"D2H pinned {mb} MB: {ms:.2f} ms  {mb / 1024 / (ms * 1e-3):.1f} GiB/s")
