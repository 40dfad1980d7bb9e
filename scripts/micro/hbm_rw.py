"""HBM bandwidth by access mix on this GPU (CUDA events, best of 5): write-only (fill), read-only (reduction), copy.
Context for the training forward kernel, whose HBM traffic is almost all writes."""
import torch
dev = 'cuda:0'
n = 1 << 30                      # 4 GiB of fp32
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)
def best(fn, bytes_moved, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return bytes_moved / (min(ts) * 1e-3) / 1e9
a.fill_(1.0); b.fill_(2.0); torch.cuda.synchronize()
print(f'write-only (fill 4 GiB):   {best(lambda: a.fill_(3.0), 4 * n):8.1f} GB/s')
print(f'write-only (zero_ 4 GiB):  {best(lambda: a.zero_(), 4 * n):8.1f} GB/s')
print(f'read-only  (sum 4 GiB):    {best(lambda: a.sum(), 4 * n):8.1f} GB/s')
print(f'copy (read + write 8 GiB): {best(lambda: b.copy_(a), 8 * n):8.1f} GB/s')
print(f'a + b -> b (12 GiB):       {best(lambda: torch.add(a, b, out=b), 12 * n):8.1f} GB/s')
