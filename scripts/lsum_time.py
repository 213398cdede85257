import torch, time, sys
sys.path.insert(0, '.')
import fruits_b200 as fruits
from fruits_b200.cache import SharedSeedCache
X = torch.randn((100000, 3, 2048), dtype=torch.float64, device="cuda").cumsum(dim=2)
for l2 in (False, True):
    for _ in range(3): out = SharedSeedCache._lsum(X, l2)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(5): out = SharedSeedCache._lsum(X, l2)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    gb = (100000 * 2048 * 8 * 2) / 1e9
    print(f"fb_lsum l2={l2}: {ms:.3f} ms, {gb / ms * 1e3:.0f} GB/s of the algorithmic {gb:.2f} GB (dim 0 read + sums written)")
