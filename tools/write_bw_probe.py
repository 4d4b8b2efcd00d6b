import torch
x = torch.empty(3, 65536*486, dtype=torch.float32, device="cuda")
def timed(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): fn(0)
    torch.cuda.synchronize(); e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e3
us = timed(lambda i: x[i%3].fill_(1.0))
print("fill 127MB: %.2f us -> %.0f GB/s" % (us, x[0].numel()*4/us/1e3))
y = torch.empty_like(x[0])
us = timed(lambda i: y.copy_(x[i%2]))
print("copy 127MB: %.2f us -> %.0f GB/s (r+w)" % (us, 2*x[0].numel()*4/us/1e3))
us = timed(lambda i: x[i%3].zero_())
print("zero (memset) 127MB: %.2f us -> %.0f GB/s" % (us, x[0].numel()*4/us/1e3))
