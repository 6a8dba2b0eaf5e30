"""Write-only HBM bandwidth next to the copy peak: is 3.4 TB/s of epilogue stores a device limit or ours?"""
import torch
n = 824 * 1024 * 1024 // 2
a = torch.empty(n, device="cuda", dtype=torch.bfloat16)
b = torch.empty(n, device="cuda", dtype=torch.bfloat16)
def t(fn, nbytes, name):
    for _ in range(3): fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for i in range(10):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
    print(f"{name:30s} {ts[5]*1e3:8.1f} us  {nbytes/ts[5]/1e6:8.1f} GB/s")
t(lambda: a.zero_(), a.numel() * 2, "zero_ (memset) 864 MB")
t(lambda: a.fill_(1.5), a.numel() * 2, "fill_ 864 MB")
t(lambda: b.copy_(a), a.numel() * 4, "copy_ 864 MB -> 864 MB")
t(lambda: torch.sum(a), a.numel() * 2, "sum (read only)")
a32 = a.view(torch.float32)
t(lambda: a32.fill_(1.5), a.numel() * 2, "fill_ f32 864 MB")
