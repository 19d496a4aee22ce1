import torch
x = torch.empty(32768*4096, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
for name, fn in (("fill_(0.5) 537MB", lambda: x.fill_(0.5)), ("zero_ 537MB", lambda: x.zero_()), ("copy_ 537MB->537MB", lambda: y.copy_(x))):
    for _ in range(5): fn()
    ts = []
    for _ in range(20):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    t = sorted(ts)[len(ts)//2]
    gb = x.numel()*4/1e9 * (2 if "copy" in name else 1)
    print(f"{name}: {t*1e3:.1f} us  {gb/(t*1e-3):.0f} GB/s")
