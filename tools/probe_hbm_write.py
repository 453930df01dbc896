"""HBM bandwidth by access mix (torch ops only; a probe, not product code): pure write (fill), pure read
(sum), copy (read+write).  Informs the roofline of write-dominated kernels (stem, shortcut convs)."""
import torch
dev = torch.device("cuda", 0)
n = 1 << 30  # 2 GiB of bf16
a = torch.empty(n, dtype=torch.bfloat16, device=dev)
b = torch.empty(n, dtype=torch.bfloat16, device=dev)
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.zero_()); print("memset (cudaMemset path) write %.0f GB/s" % (2 * n / ms / 1e6))
ms = t(lambda: a.fill_(1.5)); print("fill_ kernel write        %.0f GB/s" % (2 * n / ms / 1e6))
ms = t(lambda: b.copy_(a)); print("copy read+write           %.0f GB/s" % (4 * n / ms / 1e6))
af = a.view(torch.int16)
ms = t(lambda: af.max()); print("reduction read            %.0f GB/s" % (2 * n / ms / 1e6))
