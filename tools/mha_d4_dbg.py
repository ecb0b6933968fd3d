import sys; sys.path.insert(0,'/root/repo')
import torch, sma_b200 as S
torch.manual_seed(0)
B,L,E,heads,Skv=2,1024,32,8,1024
D=E//heads
q=torch.randn(B,L,E); k=torch.randn(B,Skv,E); v=torch.randn(B,Skv,E)
qh=q.double().view(B,L,heads,D).transpose(1,2)*(D**-0.5); kh=k.double().view(B,Skv,heads,D).transpose(1,2); vh=v.double().view(B,Skv,heads,D).transpose(1,2)
ref=(torch.softmax(qh@kh.transpose(-1,-2),-1)@vh).transpose(1,2).reshape(B,L,E)
got=S.ops.mha(q.cuda(),k.cuda(),v.cuda(),heads).cpu().double()
e=(got-ref).abs()
print('max err',float(e.max()),'mean',float(e.mean()), 'nan', int(torch.isnan(got).sum()))
print(e.view(B,L,heads,D)[0,:4,0], got.view(B,L,heads,D)[0,:2,0], ref.view(B,L,heads,D)[0,:2,0])
# per-row error pattern
er=e.view(B,L,heads,D).amax(dim=(0,2,3))
print('rows with err>1e-3:', int((er>1e-3).sum()), er[:40])
import time
qq=torch.randn(64,1024,32,device='cuda'); kk=torch.randn(64,1024,32,device='cuda'); vv=torch.randn(64,1024,32,device='cuda')
for _ in range(3): S.ops.mha(qq,kk,vv,8)
torch.cuda.synchronize(); t=time.time()
for _ in range(10): S.ops.mha(qq,kk,vv,8)
torch.cuda.synchronize(); print('ms per call', (time.time()-t)*100)
