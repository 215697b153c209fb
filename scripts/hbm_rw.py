import torch, time
x = torch.empty(300_000_000, dtype=torch.float32, device='cuda')   # 1.2 GB
y = torch.empty_like(x)
for name, fn, bytes_ in (('memset (write only)', lambda: x.zero_(), x.numel()*4),
                         ('fill 1.0 (write only)', lambda: x.fill_(1.0), x.numel()*4),
                         ('copy (read+write)', lambda: y.copy_(x), 2*x.numel()*4),
                         ('sum (read only)', lambda: x.sum(), x.numel()*4)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e)/10
    print('%-24s %.3f ms  %.2f TB/s' % (name, ms, bytes_/ms/1e9))
