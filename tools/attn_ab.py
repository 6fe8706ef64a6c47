import sys, os, json, torch
sys.path.insert(0, os.getcwd())
from maest_b200 import ops
B, N = 64, 1685
qkv = torch.randn(B*N, 2304, device="cuda").half()
def timeit(fn, n=8):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(True),torch.cuda.Event(True); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
q,k,v = qkv[:1685].view(1,1685,3,12,64).permute(2,0,3,1,4).double()
ref = (torch.softmax((q@k.transpose(-1,-2))*0.125,-1)@v).transpose(1,2).reshape(1685,768)
o = ops.attention(qkv[:1685].contiguous(), 1, 1685, 12, 0)
rel = float((o.double()-ref).norm()/ref.norm())
print(json.dumps(dict(lib=os.environ.get("MAEST_B200_LIB","default")[-12:], ms=timeit(lambda: ops.attention(qkv,B,N,12,0)), rel=rel)))
