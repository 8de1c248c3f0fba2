import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
sys.path.insert(0, '/root/repo/tests')
import pytest
from tests import test_gpu_diffsample as T
from ratspn_b200 import diffsample

class MP:
    def __init__(self): self.saved=None
    def setattr(self, mod, name, val):
        self.saved=(mod,name,getattr(mod,name)); setattr(mod,name,val)
    def undo(self):
        if self.saved: setattr(*self.saved); self.saved=None

def report(tag, m, **kw):
    mp = MP()
    li = kw.get('layer_index', m.max_layer_index)
    lead = kw.pop('lead')
    a = T._run(m, False, mp, noise=T._noise(m, li, lead), **kw)
    b = T._run(m, True, mp, noise=T._noise(m, li, lead), **kw)
    mp.undo()
    print(tag, 'sample err', float((a[0]-b[0]).abs().max()))
    for k in b[1]:
        ga = a[1].get(k)
        if ga is None: print('   ', k, 'MISSING'); continue
        print('   %-40s err %.3e scale %.3e' % (k, float((ga-b[1][k]).abs().max()), float(b[1][k].abs().max())))

torch.set_printoptions(precision=4, linewidth=200)
for arch in [(10,3,1,2,2,1),(14,3,1,2,2,1)]:
    m = T._ratspn(*arch)
    print(arch, 'card', m._leaf.cardinality, 'leaf out', m._leaf.out_features if hasattr(m._leaf,'out_features') else None, [ (type(l).__name__, l.in_features, getattr(l,'_pad',None)) for l in m._inner_layers])
    res = []
    for comp in (False, True):
        mp = MP()
        if comp:
            def wrapped(spn, ctx, class_index, evidence, layer_index, noise, fuse_post=False):
                return diffsample.sample_onehot_composite(spn, ctx, class_index, evidence, layer_index, noise), False
            mp.setattr(diffsample, "sample_onehot", wrapped)
        m.zero_grad(set_to_none=True)
        ctx = m.sample(mode="onehot", n=1, layer_index=2, do_sample_postprocessing=False, noise=T._noise(m, 2, (2,1,1)))
        y = ctx.sample
        y.sum().backward()
        res.append(m._inner_layers[1].weight_param.grad.clone())
        mp.undo()
    print('fused', res[0][0,:,:,:,0].permute(0,2,1))
    print('comp ', res[1][0,:,:,:,0].permute(0,2,1))
