"""Read-only HBM ceiling on this box (development aid): torch.sum over the stencil benchmark's input, against which the
reduction stencils (action / topological charge: 16 B read per site, nothing written) are to be judged."""
import torch
x = torch.rand(49152, 2, 32, 32, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
def t(fn, n=9):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[n // 2]
nb = x.numel() * 8
print(f"torch.sum(x)            {nb / t(lambda: torch.sum(x)) / 1e6:8.0f} GB/s read")
print(f"torch.sum(x, dim=(1,2,3)) {nb / t(lambda: torch.sum(x, dim=(1, 2, 3))) / 1e6:6.0f} GB/s read")
print(f"y.copy_(x)              {2 * nb / t(lambda: y.copy_(x)) / 1e6:8.0f} GB/s read+write")
print(f"x.view(-1)[::2].sum()   n/a")
