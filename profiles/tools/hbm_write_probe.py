"""Write-only / read-only / copy bandwidth of this B200 for buffers of the C2 output size (288 MB) and larger."""
import torch
for mb in (288, 1152):
    n = mb * 1000 * 1000 // 4
    a = torch.empty(n, dtype=torch.float32, device="cuda")
    b = torch.empty(n, dtype=torch.float32, device="cuda")
    def t(fn, reps=20):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    w = t(lambda: a.fill_(1.0)); z = t(lambda: a.zero_()); c = t(lambda: b.copy_(a)); r = t(lambda: a.sum())
    print(f"{mb} MB: fill {mb/w:.0f} GB/s  memset {mb/z:.0f} GB/s  copy(read+write) {2*mb/c:.0f} GB/s  read(sum) {mb/r:.0f} GB/s")
