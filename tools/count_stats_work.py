import ctypes as C, sys, numpy as np
sys.path.insert(0,'/root/repo')
import torch
from bench import WORKLOADS, make_model, make_batch
from htk_b200.estep import ForwardBackward
from htk_b200 import capi
cfg=WORKLOADS[sys.argv[1] if len(sys.argv)>1 else 'cfg3']
fm=make_model(cfg); dev=torch.device('cuda',0)
fb=ForwardBackward(fm)
b,df=make_batch(fm,cfg,256,1000,dev)
fb.FBFile(b, device_feat_ptr=df.data_ptr())
out=(C.c_ulonglong*8)()
capi.load().hfbgpu_debug_counters(out)
inb,valid,pairs,act,chunks=out[0],out[1],out[2],out[3],out[4]
fr=256*cfg['T']
print('per frame: in-beam (pos,frame) %.2f, after skip %.2f, (mix,frame) pairs %.1f, active mixtures per chunk %.2f, chunks per frame %.3f'%(inb/fr, valid/fr, pairs/fr, act/max(chunks,1), chunks/fr))
