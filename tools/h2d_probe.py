"""H2D / D2H bandwidth of page-locked torch tensors of the bench's sizes (diagnostic)."""
import numpy as np
import torch

n = 16_000_000
dev = torch.device("cuda:0")
for name, shape, dt in [("pos f32[n,2]", (n, 2), torch.float32), ("q f32[n]", (n,), torch.float32),
                        ("q via numpy pin", None, None), ("orig i32[n]", (n,), torch.int32)]:
    if shape is None:
        h = torch.from_numpy(np.ascontiguousarray(np.random.rand(n).astype(np.float32))).pin_memory()
    else:
        h = torch.empty(shape, dtype=dt).pin_memory()
        h.zero_()
    d = torch.empty_like(h, device=dev)
    for direction in ("h2d", "d2h"):
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            if direction == "h2d":
                d.copy_(h, non_blocking=True)
            else:
                h.copy_(d, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        mb = h.numel() * h.element_size() / 1e6
        print(f"{name:18s} {direction} {mb:7.1f} MB  {min(ts):6.2f} ms  {mb / min(ts):6.1f} GB/s  pinned={h.is_pinned()}")
