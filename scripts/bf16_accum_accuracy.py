import sys; sys.path.insert(0,'/root/repo')
import torch, numpy as np
from pctrans_b200 import _cabi, synth
from oracle import msda_oracle as oracle
dev='cuda:0'
for regime,seed in (("R",0),("U",1)):
    value, sh, lsi, loc, attn = synth.make_inputs(2, synth.CVPPP_512, M=8, D=32, P=4, regime=regime, seed=seed)
    g = torch.randn(2, value.shape[1], 256, generator=torch.Generator().manual_seed(1))
    vb, gb = value.bfloat16(), g.bfloat16()
    ref = oracle.backward_c(vb.float(), sh, lsi, loc, attn, gb.float())[0]
    for flags in (0, 8):
        gv = _cabi.backward(vb.to(dev), sh.to(dev), lsi.to(dev), loc.to(dev), attn.to(dev), gb.to(dev), flags=flags)[0].float().cpu().numpy()
        err = np.abs(gv-ref)
        print(regime, "flags",flags, "max-rel-to-max", err.max()/np.abs(ref).max(), "rms-rel", np.sqrt((err**2).mean())/np.sqrt((ref**2).mean()))
