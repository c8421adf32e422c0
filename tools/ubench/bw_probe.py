import torch
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); y = torch.empty_like(x)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best * 1e-3
print("memset 1GiB  GB/s:", (1 << 30) / t(lambda: x.zero_()) / 1e9)
print("fill bf16    GB/s:", (1 << 30) / t(lambda: x.view(torch.bfloat16).fill_(1.0)) / 1e9)
print("copy 1GiB (r+w) GB/s:", 2 * (1 << 30) / t(lambda: y.copy_(x)) / 1e9)
print("read (sum int32) GB/s:", (1 << 30) / t(lambda: x.view(torch.int32).sum()) / 1e9)
