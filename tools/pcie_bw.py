import torch, time, subprocess
print(subprocess.run(["nvidia-smi","--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current","--format=csv"],capture_output=True,text=True).stdout)
n=722*1000*1000
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda')
h2=torch.empty(n//3,dtype=torch.uint8).pin_memory(); d2=torch.empty(n//3,dtype=torch.uint8,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
for it in range(3):
    torch.cuda.synchronize(); t=time.perf_counter()
    d.copy_(h,non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("H2D %.1f GB/s"%(n/dt/1e9))
for it in range(2):
    torch.cuda.synchronize(); t=time.perf_counter()
    h2.copy_(d2,non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("D2H %.1f GB/s"%(n/3/dt/1e9))
for it in range(2):
    torch.cuda.synchronize(); t=time.perf_counter()
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("bidir: H2D %.1f GB/s + D2H %.1f GB/s in %.1f ms"%(n/dt/1e9, n/3/dt/1e9, dt*1e3))
# chunked copies of 46 MB like the library does
c=64*360960
for it in range(2):
    torch.cuda.synchronize(); t=time.perf_counter()
    for k in range(0,n-c,c): d[k:k+c].copy_(h[k:k+c],non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("H2D 23MB chunks %.1f GB/s"%(((n-c)//c*c)/dt/1e9))
