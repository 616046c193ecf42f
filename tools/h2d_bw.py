"""Host-to-device bandwidth of this box for the e2e leg of bench.py: one pinned 300 MB buffer (a 1.25e7 x 6 float32
candidate shard), whole and in 25 MB / 50 MB pieces."""
import torch

n = 12_500_000 * 6
host = torch.empty(n, dtype=torch.float32).pin_memory()
dev = torch.empty(n, dtype=torch.float32, device="cuda")
for piece in (n, (1 << 20) * 6, (1 << 21) * 6):
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(3):
        for o in range(0, n, piece):
            dev[o:o + piece].copy_(host[o:o + piece], non_blocking=True)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 3
    print(f"pieces of {piece * 4 / 1e6:7.1f} MB: {n * 4 / ms / 1e6:6.1f} GB/s ({ms:.2f} ms per 300 MB)")
