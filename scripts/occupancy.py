import sys, ctypes as C
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[R, os.path.join(R,'neuralnet-tracker-traincode_b200')]
import torch
torch.cuda.init(); torch.zeros(1,device='cuda')
from trackertraincode_b200 import _native as N
for cl in (1,2,4,8):
    a=C.c_int(0); b=C.c_int(0)
    rc=N.lib.b200aug_fused_occupancy(129,129,0,cl,C.byref(a),C.byref(b))
    print("cluster",cl,"rc",rc,"ctas/SM",a.value,"active clusters",b.value,"-> CTAs",b.value*cl, "smem", N.lib.b200aug_fused_smem_bytes(129,129,0))
