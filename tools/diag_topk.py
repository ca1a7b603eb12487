import sys
sys.path.insert(0,'pytorch-human-pose_b200'); sys.path.insert(0,'.')
import torch, numpy as np
from hpdecode import synth_net, synth, ops
def run(scale, name):
    B=scale['hm_lo'].shape[0]; S=512
    bufs=ops.DecodeBuffers(B,17,S,S,2,30,'cuda:0'); p=ops.make_params(B,17,S,S,2,30,0.05,0.5)
    ops.run_stage('aggregate_nms',bufs,p,scales=[scale]); torch.cuda.synchronize()
    for force in (0,1):
        p.force_generic=force
        ops.run_stage('topk',bufs,p); torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); ops.run_stage('topk',bufs,p); e1.record(); torch.cuda.synchronize()
        print(name,'force_exact' if force else 'fast', 'topk ms %.3f'%e0.elapsed_time(e1))
    mask=bufs.nms_mask.cpu().numpy().view(np.uint32); hm=bufs.agg_hm.cpu().numpy()
    bits=((mask[...,None]>>np.arange(32,dtype=np.uint32))&1).reshape(B,17,S,S).astype(bool)
    surv=bits.sum((2,3)); pos=(bits&(hm>0)).sum((2,3))
    sk=bufs.scores_k.cpu().numpy()
    print(name,'survivors/row mean %.0f max %d; positive survivors mean %.0f; rows with <30 positives: %d/%d; rows with dup top-30 values: %d'%(surv.mean(),surv.max(),pos.mean(),(pos<30).sum(),pos.size, sum(len(np.unique(r))<30 for r in sk.reshape(-1,30))))
    wm=bufs.nms_wmax.cpu().numpy(); print(name,'words with wmax>0 per row: %.0f of %d'%((wm>0).sum((2,3)).mean(), wm.shape[2]*wm.shape[3]))
run(synth_net.network_outputs(8,512,True,seed=1,device='cuda:0'),'hrnet')
run({k:torch.from_numpy(v).cuda() for k,v in synth.netlike(8,512,True,seed=1)[0].items()},'netlike')
