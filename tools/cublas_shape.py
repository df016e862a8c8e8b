import torch, time
dev="cuda"
for (M,N,K) in ((20000,131072,1536),(20000,262144,1536),(8192,8192,8192)):
    a=torch.randn(M,K,device=dev,dtype=torch.bfloat16); b=torch.randn(N,K,device=dev,dtype=torch.bfloat16)
    out=torch.empty(M,N,device=dev,dtype=torch.bfloat16)
    for _ in range(3): torch.matmul(a,b.t(),out=out)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    iters=max(3,int(2e15/(2*M*N*K)))   # ~2 PFLOP of work: long enough to hit the power cap
    e0.record()
    for _ in range(iters): torch.matmul(a,b.t(),out=out)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/iters
    print("cuBLAS bf16 %d x %d x %d: %.3f ms -> %.1f TFLOP/s (%d iters, %.1f s)"%(M,N,K,ms,2*M*N*K/ms*1e-9,iters,ms*iters*1e-3))
    del a,b,out
