import torch, sys, os, json
sys.path.insert(0, os.getcwd())
from pikazoo_b200 import _lib
L=_lib.load()
nbytes = 280<<20
nbytes -= nbytes % (8960*4)
bufs=[torch.empty(nbytes//4, dtype=torch.int32, device='cuda') for _ in range(2)]
s_=torch.cuda.current_stream().cuda_stream
out={}
for mode in range(5):
    for k in range(4): _lib.check(L.pz_probe_write(bufs[k&1].data_ptr(), nbytes, mode, s_))
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(40): _lib.check(L.pz_probe_write(bufs[k&1].data_ptr(), nbytes, mode, s_))
    b.record(); torch.cuda.synchronize()
    out[mode]=nbytes*40/(a.elapsed_time(b)*1e-3)/1e9
a.record()
for k in range(40): bufs[k&1].zero_()
b.record(); torch.cuda.synchronize()
out['zero']=nbytes*40/(a.elapsed_time(b)*1e-3)/1e9
print(json.dumps(out))
