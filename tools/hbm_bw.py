import torch
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
def t(f, n=10):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.fill_(1)); print(f"fill 1 GiB: {ms:.3f} ms = {1.0737/ms:.2f} TB/s write")
ms = t(lambda: y.copy_(x)); print(f"copy 1 GiB: {ms:.3f} ms = {2*1.0737/ms:.2f} TB/s r+w")
ms = t(lambda: x.sum()); print(f"sum 1 GiB (u8): {ms:.3f} ms = {1.0737/ms:.2f} TB/s read")
xf = x.view(torch.float32)
ms = t(lambda: xf.sum()); print(f"sum 1 GiB (f32): {ms:.3f} ms = {1.0737/ms:.2f} TB/s read")
