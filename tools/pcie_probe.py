import torch, time
n = 512*1024*1024
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device='cuda')
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(f, reps=5):
    f(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
def h2d(): d.copy_(h, non_blocking=True)
def d2h(): h2.copy_(d2, non_blocking=True)
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print('H2D GB/s', n/t(h2d)/1e9); print('D2H GB/s', n/t(d2h)/1e9); print('both each GB/s', n/t(both)/1e9)
# chunked 64MB
c=64*1024*1024
def h2d_chunks():
    for i in range(0,n,c): d[i:i+c].copy_(h[i:i+c], non_blocking=True)
print('H2D 64MB chunks GB/s', n/t(h2d_chunks)/1e9)
# pageable
hp = torch.empty(n, dtype=torch.uint8)
def h2d_pageable(): d.copy_(hp)
print('H2D pageable GB/s', n/t(h2d_pageable,2)/1e9)
