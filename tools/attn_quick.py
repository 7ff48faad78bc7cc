import torch, sys
sys.path.insert(0, '.')
from torchacc_b200.ops import attention as A
A.set_attention_backend("native")
q=(torch.randn(1,512,4,128,device='cuda')*0.8).bfloat16(); k=(torch.randn(1,512,2,128,device='cuda')*0.8).bfloat16(); v=(torch.randn(1,512,2,128,device='cuda')*0.8).bfloat16()
o=A.flash_attn_func(q,k,v,causal=True); torch.cuda.synchronize()
r,_=A.attention_reference(q.float(),k.float(),v.float(),None,True,(-1,-1))
print("quick ok", float((o.float()-r).abs().max()))
