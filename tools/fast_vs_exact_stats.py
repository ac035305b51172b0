import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/optical-flow-fpga_b200')
import of_b200, synthetic
from oracle import lk_float_oracle as orc
for (H,W,L,I) in ((480,640,5,10),(480,640,3,3),(1080,1920,5,10)):
    prev,curr,_=synthetic.make_pairs_numpy(1,H,W,seed=5)
    p,c=prev[0],curr[0]
    ue,ve=of_b200.lk_pyramidal(p,c,L,5,I,mode=0)
    uf,vf=of_b200.lk_pyramidal(p,c,L,5,I,mode=1)
    if H<=480:
        uo,vo=orc.lucas_kanade_pyramidal(p,c,L,5,I)
        print(H,W,L,I,'exact==oracle',np.array_equal(ue.view(np.uint32),uo.view(np.uint32)))
    d=np.maximum(np.abs(uf-ue),np.abs(vf-ve))
    print(H,W,L,I,'fast vs exact: max',d.max(),'frac>1e-3',(d>1e-3).mean(),'mean',d.mean(),'median',np.median(d), 'mean|u|',np.abs(ue).mean())
