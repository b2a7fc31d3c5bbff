"""What the NVLink path delivers on this box, as a yardstick for the exchange kernel (K10): bidirectional peer copies
between GPU 0 and GPU 1 with the copy engines (torch copy_ on two streams) for several sizes."""
import torch

assert torch.cuda.device_count() >= 2
a0 = torch.empty(1 << 30, dtype=torch.float32, device="cuda:0")   # 4 GiB
a1 = torch.empty(1 << 30, dtype=torch.float32, device="cuda:1")
b0 = torch.empty_like(a0)
b1 = torch.empty_like(a1)
s0 = torch.cuda.Stream(device="cuda:0")
s1 = torch.cuda.Stream(device="cuda:1")
for both in (False, True):
    for _ in range(2):
        with torch.cuda.stream(s0):
            b1.copy_(a0, non_blocking=True)          # 0 -> 1, issued by GPU 0
        if both:
            with torch.cuda.stream(s1):
                b0.copy_(a1, non_blocking=True)      # 1 -> 0, issued by GPU 1
        torch.cuda.synchronize("cuda:0"); torch.cuda.synchronize("cuda:1")
    with torch.cuda.device(0):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    with torch.cuda.stream(s0):
        e0.record(s0)
        for _ in range(reps):
            b1.copy_(a0, non_blocking=True)
        e1.record(s0)
    if both:
        with torch.cuda.stream(s1):
            for _ in range(reps):
                b0.copy_(a1, non_blocking=True)
    torch.cuda.synchronize("cuda:0"); torch.cuda.synchronize("cuda:1")
    ms = e0.elapsed_time(e1) / reps
    print(f"{'bidirectional' if both else 'one direction'}: 4 GiB peer copy 0 -> 1 in {ms:.3f} ms = {a0.numel() * 4 / ms / 1e6:.0f} GB/s per direction")
