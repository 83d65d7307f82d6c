import sys, torch, time
sys.path.insert(0, '/root/repo')
from va_vae_b200 import engine as E
E.set_gemm_two_cta(True)
dev = torch.device('cuda', 0)
def rel(a, b): return float((a.double()-b.double()).norm()/b.double().norm())
for (M,N,K) in [(256,256,64),(256,256,128),(512,512,256),(1024,3456,1152),(4096,1152,3072),(1000,384,384),(16384,6144,1152),(384,256,128)]:
    g = torch.Generator().manual_seed(M+N+K)
    A = (torch.randn(M,K,generator=g)*0.5).to(dev).bfloat16().contiguous()
    W = (torch.randn(N,K,generator=g)*0.5).to(dev).bfloat16().contiguous()
    b = torch.randn(N,generator=g).to(dev)
    out = E.test_gemm_bf16(A,W,b); torch.cuda.synchronize()
    ref = A.float()@W.float().t()+b
    print(M,N,K,'rel',rel(out,ref), flush=True)
# timing
for two in (False, True):
    E.set_gemm_two_cta(two)
    M,N,K = 131072, 4096, 768
    A = torch.randn(M,K,device=dev).bfloat16(); W = torch.randn(N,K,device=dev).bfloat16(); b = torch.zeros(N,device=dev)
    for _ in range(3): E.test_gemm_bf16(A,W,b)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): E.test_gemm_bf16(A,W,b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    print('two_cta',two,'ms',ms,'TFLOP/s',2*M*N*K/ms/1e9, flush=True)
