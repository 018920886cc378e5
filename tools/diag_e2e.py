import sys, os; sys.path.insert(0, os.getcwd()); sys.path.insert(0,"tests")
import numpy as np, torch
from gpu_helpers import *
from helpers import *
from oracle import pan as op
for cname,B in [("C2",24),("C5",8),("C4",24)]:
    cfg=CONFIGS[cname]
    inp=make_inputs(cfg,B=B,scene="obstacles")
    pan=make_pan(cfg,K=2,max_envs=B)
    S,U,D,md=run_pan(pan,inp)
    mk=oracle_factory(cfg,K=2)
    for b in range(B):
        o=mk(); vel=None if inp["velocities"] is None else inp["velocities"][b]
        So,Uo,Do=o.forward(inp["nom_s"][b],inp["nom_u"][b],inp["ref_s"][b],inp["ref_us"][b],inp["points"][b],vel,keep_trace=True)
        e=max(rel_err(S[b],So),rel_err(U[b],Uo),rel_err(D[b],Do[0]))
        if e>1e-4:
            # K=1 check from same start
            p1=make_pan(cfg,K=1,max_envs=1)
            one={k:(None if v is None else v[b:b+1]) for k,v in inp.items()}
            S1,U1,D1,md1=run_pan(p1,one)
            t0=o.trace[0]
            e1=max(rel_err(S1[0],t0["S"]),rel_err(U1[0],t0["U"]),rel_err(D1[0],t0["D"][0]))
            # second iteration teacher-forced from oracle's first
            one2=dict(one,nom_s=t0["S"][None],nom_u=t0["U"][None])
            S2,U2,D2,md2=run_pan(make_pan(cfg,K=1,max_envs=1),one2)
            e2=max(rel_err(S2[0],So),rel_err(U2[0],Uo),rel_err(D2[0],Do[0]))
            print(cname,b,"K=2 err %.2e | iter1 err %.2e | iter2 teacher-forced err %.2e | dU max %.2e"%(e,e1,e2,np.abs(U[b]-Uo).max()))
