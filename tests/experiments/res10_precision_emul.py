#!/usr/bin/env python
"""CPU emulation (oracle-side experiment, not a test) behind precision="split_act": the 10-block residual net with
fp16 / exact activations and weights rounded to fp16 by round-to-nearest or by error diffusion along K, against the
fp64 oracle.  Round-1 result (48 positions, max |d log p|): fp16 x fp16 2.1e-3; exact activations x RN weights 1.16e-3;
exact activations x diffusion-rounded weights 4.7e-4; fp16 activations x exact weights 1.6e-3.
    python tests/experiments/res10_precision_emul.py"""
import os
import sys
import numpy as np
import torch
import torch.nn.functional as F
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import net as onet
from helpers import oboard_from, synth_position
torch.set_num_threads(8)
W=H=15; NB=10
arg,aux=onet.init_params("resnet",W,H,seed=0,n_blocks=NB,synthetic_stats=True)
P={k:torch.as_tensor(v) for k,v in list(arg.items())+list(aux.items())}
boards=[oboard_from(W,H,5,synth_position(W,H,5,1234+i)) for i in range(48)]
st=np.stack([np.ascontiguousarray(b.current_state(),dtype=np.float32) for b in boards])
calib=[oboard_from(W,H,5,synth_position(W,H,5,5000+i)) for i in range(32)]
stc=np.stack([np.ascontiguousarray(b.current_state(),dtype=np.float32) for b in calib])
rp,rv,rl=onet.forward(arg,aux,st,"resnet",n_blocks=NB,dtype=torch.float64,return_logits=True)

def h16(t): return t.half().float()
def fold(wn,bn_prefix,fix_gamma,mean_key,var_key):
    w=P[wn+"_weight"]; b=P[wn+"_bias"]
    g=torch.ones_like(P[bn_prefix+"_gamma"]) if fix_gamma else P[bn_prefix+"_gamma"]
    s=g/torch.sqrt(P[var_key]+onet.BN_EPS)
    return w*s[:,None,None,None], (b-P[mean_key])*s+P[bn_prefix+"_beta"]

def diffuse(w, mu=None):
    """sigma-delta rounding to fp16 along the (cin,kh,kw) axis of each output channel, error weighted by mu[cin]"""
    O=w.shape[0]
    wf=w.reshape(O,-1).double()
    K=wf.shape[1]
    if mu is None: m=torch.ones(K,dtype=torch.float64)
    else: m=mu.double().repeat_interleave(w.shape[2]*w.shape[3]).clamp_min(1e-6)
    out=torch.empty_like(wf)
    acc=torch.zeros(O,dtype=torch.float64)
    # candidates: round down / up in fp16
    rn=wf.float().half()
    for k in range(K):
        x=wf[:,k]
        r=rn[:,k].double()
        # neighbours
        up=torch.nextafter(rn[:,k],torch.tensor(float('inf'),dtype=torch.half)).double()
        dn=torch.nextafter(rn[:,k],torch.tensor(float('-inf'),dtype=torch.half)).double()
        lo=torch.where(r<=x,r,dn); hi=torch.where(r<=x,up,r)
        e_lo=(lo-x)*m[k]; e_hi=(hi-x)*m[k]
        pick_hi=(acc+e_hi).abs()<(acc+e_lo).abs()
        q=torch.where(pick_hi,hi,lo)
        acc=acc+(q-x)*m[k]
        out[:,k]=q
    return out.reshape(w.shape).float()

def run(states, act_round, wmode, mus=None, collect=False):
    x=torch.as_tensor(states)
    means=[]
    def conv(x,wn,bnp,fix,mk,vk,idx,relu=True,resid=None):
        w,sh=fold(wn,bnp,fix,mk,vk)
        if collect: means.append(x.mean(dim=(0,2,3)))
        if wmode=="rn": wq=h16(w)
        elif wmode=="exact": wq=w
        elif wmode=="diff": wq=diffuse(w)
        elif wmode=="diffmu": wq=diffuse(w,mus[idx])
        y=F.conv2d(x,wq,None,padding=1)+sh[None,:,None,None]
        if resid is not None: y=y+resid
        if relu: y=F.relu(y)
        return act_round(y)
    i=0
    x=act_round(x)
    x=conv(x,"res_conv1","res_conv1",True,"res_conv1_mean","res_conv1_var",i); i+=1
    for b in range(1,NB+1):
        idn=x
        y=conv(x,"convA%d"%b,"bnA%d"%b,False,"bnA%d_moving_mean"%b,"bnA%d_moving_var"%b,i); i+=1
        x=conv(y,"convB%d"%b,"bnB%d"%b,False,"bnB%d_moving_mean"%b,"bnB%d_moving_var"%b,i,resid=idn); i+=1
    # heads in fp32 (device heads are near-fp32)
    def ca(x,name):
        y=F.conv2d(x,P[name+"_weight"],P[name+"_bias"])
        y=onet._bn(y,P[name+"_gamma"],P[name+"_beta"],P[name+"_mean"],P[name+"_var"],True)
        return F.relu(y)
    B=x.shape[0]
    p=ca(x,"conv3_1_1").reshape(B,-1); logits=p@P["fc_3_1_1_weight"].t()+P["fc_3_1_1_bias"]
    v=torch.tanh(ca(x,"conv3_2_1").reshape(B,-1)@P["fc_3_2_1_weight"].t()+P["fc_3_2_1_bias"])
    return torch.log_softmax(logits,1).numpy(), v.numpy(), means

ident=lambda t:t
rlogp=np.log(rp)
def report(name,lp,v):
    print("%-34s max|dlogp| %.2e  mean %.2e  max|dv| %.2e"%(name,np.abs(lp-rlogp).max(),np.abs(lp-rlogp).mean(),np.abs(v-rv).max()))
with torch.no_grad():
    lp,v,_=run(st,ident,"exact"); report("fp32 emulation (sanity)",lp,v)
    lp,v,_=run(st,h16,"rn"); report("fp16 act + fp16 w (RN)",lp,v)
    lp,v,_=run(st,ident,"rn"); report("exact act + fp16 w (RN)",lp,v)
    lp,v,_=run(st,h16,"exact"); report("fp16 act + exact w",lp,v)
    lp,v,_=run(st,ident,"diff"); report("exact act + fp16 w (diffusion)",lp,v)
    _,_,mus=run(stc,ident,"exact",collect=True)
    lp,v,_=run(st,ident,"diffmu",mus=mus); report("exact act + fp16 w (mu-diffusion)",lp,v)
    lp,v,_=run(st,h16,"diffmu",mus=mus); report("fp16 act + fp16 w (mu-diffusion)",lp,v)
