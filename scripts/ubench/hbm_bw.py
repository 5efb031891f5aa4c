import torch
dev = torch.device("cuda:0")
x = torch.empty(1 << 30, dtype=torch.float32, device=dev)   # 4 GB
y = torch.empty(1 << 30, dtype=torch.float32, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: x.zero_()); print("write-only  %.1f GB/s" % (4.295 / ms * 1e3))
ms = t(lambda: x.sum());   print("read-only   %.1f GB/s" % (4.295 / ms * 1e3))
ms = t(lambda: y.copy_(x)); print("copy        %.1f GB/s" % (2 * 4.295 / ms * 1e3))
