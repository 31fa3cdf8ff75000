# Dev probe: pure-write, pure-read and copy bandwidth of the box (torch kernels), to put the
# write-only covariance build in context.
import torch
dev = "cuda"
n = 1 << 30          # 8 GiB of doubles
x = torch.empty(n, dtype=torch.float64, device=dev)
y = torch.empty(n, dtype=torch.float64, device=dev)
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
ms = t(lambda: x.zero_());            print("zero_  (memset)     : %.0f GB/s written" % (n * 8 / ms / 1e6))
ms = t(lambda: x.fill_(1.5));         print("fill_  (store kernel): %.0f GB/s written" % (n * 8 / ms / 1e6))
ms = t(lambda: y.copy_(x));           print("copy_               : %.0f GB/s read + written" % (2 * n * 8 / ms / 1e6))
ms = t(lambda: x.sum());              print("sum    (read kernel) : %.0f GB/s read" % (n * 8 / ms / 1e6))
ms = t(lambda: torch.add(x, 1.0, out=y)); print("add out=            : %.0f GB/s read + written" % (2 * n * 8 / ms / 1e6))
