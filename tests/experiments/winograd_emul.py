#!/usr/bin/env python
"""CPU emulation (oracle-side experiment, not a test): the 6-conv reference net with Winograd F(2x2,3x3) convolutions
whose transformed inputs and weights are rounded to fp16 (fp32 accumulation), against the fp64 oracle.
Result on 64 SURVEY 8(d) positions (round 1), max |d log p|: direct fp16 operands 2.5e-4, F(2x2,3x3) with fp16
operands 4.1e-4, F(4x4,3x3) 6.0e-4 - all inside the 1e-3 parity budget, i.e. the 2.25x / 4x MAC reductions are
numerically available to the conv kernels.
    python tests/experiments/winograd_emul.py"""
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
W=H=15
arg,aux=onet.init_params("simple",W,H,seed=0,synthetic_stats=True)
P={k:torch.as_tensor(v) for k,v in list(arg.items())+list(aux.items())}
boards=[oboard_from(W,H,5,synth_position(W,H,5,1234+i)) for i in range(64)]
st=np.stack([np.ascontiguousarray(b.current_state(),dtype=np.float32) for b in boards])
rp,rv=onet.forward(arg,aux,st,"simple",dtype=torch.float64)
rlogp=np.log(rp)
h16=lambda t:t.half().float()
BT=torch.tensor([[1,0,-1,0],[0,1,1,0],[0,-1,1,0],[0,1,0,-1]],dtype=torch.float32)
G=torch.tensor([[1,0,0],[.5,.5,.5],[.5,-.5,.5],[0,0,1]],dtype=torch.float32)
AT=torch.tensor([[1,1,1,0],[0,1,-1,-1]],dtype=torch.float32)
def fold(name):
    w=P[name+"_weight"]; b=P[name+"_bias"]
    s=1.0/torch.sqrt(P[name+"_var"]+onet.BN_EPS)
    return w*s[:,None,None,None], (b-P[name+"_mean"])*s+P[name+"_beta"]
def conv_direct(x,w,round_w=True):
    return F.conv2d(x,h16(w) if round_w else w,None,padding=1)
def conv_wino(x,w,round_ops=True):
    B_,C,_,_=x.shape
    xp=F.pad(x,(1,2,1,2))                       # 18x18: tiles cover outputs 0..15
    patches=xp.unfold(2,4,2).unfold(3,4,2)      # B,C,8,8,4,4
    V=torch.einsum('ij,bcxyjk,lk->bcxyil',BT,patches,BT)
    U=torch.einsum('ij,ocjk,lk->ocil',G,w,G)
    if round_ops: V=h16(V); U=h16(U)
    M=torch.einsum('bcxyil,ocil->boxyil',V,U)
    Y=torch.einsum('ij,boxyjk,lk->boxyil',AT,M,AT)   # B,O,8,8,2,2
    Y=Y.permute(0,1,2,4,3,5).reshape(B_,-1,16,16)[:,:,:15,:15]
    return Y
BT4=torch.tensor([[4,0,-5,0,1,0],[0,-4,-4,1,1,0],[0,4,-4,-1,1,0],[0,-2,-1,2,1,0],[0,2,-1,-2,1,0],[0,4,0,-5,0,1]],dtype=torch.float32)
G4=torch.tensor([[1/4,0,0],[-1/6,-1/6,-1/6],[-1/6,1/6,-1/6],[1/24,1/12,1/6],[1/24,-1/12,1/6],[0,0,1]],dtype=torch.float32)
AT4=torch.tensor([[1,1,1,1,1,0],[0,1,-1,2,-2,0],[0,1,1,4,4,0],[0,1,-1,8,-8,1]],dtype=torch.float32)
def conv_wino4(x,w,round_ops=True):
    """F(4x4,3x3) (Lavin & Gray): 6x6 input tiles, 36 products per 16 outputs = 4x fewer MACs than direct"""
    B_,C,_,_=x.shape
    xp=F.pad(x,(1,2,1,2))                       # 18x18: 4x4 tiles of 4 outputs cover 0..15
    patches=xp.unfold(2,6,4).unfold(3,6,4)      # B,C,4,4,6,6
    V=torch.einsum('ij,bcxyjk,lk->bcxyil',BT4,patches,BT4)
    U=torch.einsum('ij,ocjk,lk->ocil',G4,w,G4)
    if round_ops: V=h16(V); U=h16(U)
    M=torch.einsum('bcxyil,ocil->boxyil',V,U)
    Y=torch.einsum('ij,boxyjk,lk->boxyil',AT4,M,AT4)   # B,O,4,4,4,4
    return Y.permute(0,1,2,4,3,5).reshape(B_,-1,16,16)[:,:,:15,:15]
def run(mode):
    x=h16(torch.as_tensor(st))
    for name,_ in onet.SIMPLE_TRUNK:
        w,sh=fold(name)
        if mode=="fp32": y=F.conv2d(x,w,None,padding=1)
        elif mode=="direct16": y=conv_direct(x,w)
        elif mode=="wino16": y=conv_wino(x,w)
        elif mode=="wino32": y=conv_wino(x,w,round_ops=False)
        elif mode=="wino4_16": y=conv_wino4(x,w)
        elif mode=="wino4_32": y=conv_wino4(x,w,round_ops=False)
        x=F.relu(y+sh[None,:,None,None])
        if mode!="fp32": x=h16(x)
    def ca(x,name):
        y=F.conv2d(x,P[name+"_weight"],P[name+"_bias"])
        y=onet._bn(y,P[name+"_gamma"],P[name+"_beta"],P[name+"_mean"],P[name+"_var"],True)
        return F.relu(y)
    B_=x.shape[0]
    p=ca(x,"conv3_1_1").reshape(B_,-1); logits=p@P["fc_3_1_1_weight"].t()+P["fc_3_1_1_bias"]
    v=torch.tanh(ca(x,"conv3_2_1").reshape(B_,-1)@P["fc_3_2_1_weight"].t()+P["fc_3_2_1_bias"])
    return torch.log_softmax(logits,1).numpy(), v.numpy()
with torch.no_grad():
    for mode in ("fp32","wino32","direct16","wino16","wino4_32","wino4_16"):
        lp,v=run(mode)
        print("%-9s max|dlogp| %.2e mean %.2e max|dv| %.2e"%(mode,np.abs(lp-rlogp).max(),np.abs(lp-rlogp).mean(),np.abs(v-rv).max()))
