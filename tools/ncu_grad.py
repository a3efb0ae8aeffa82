"""Small runs of K2 (implicit gradient) and K3/K4/K4b (BPTT) for ncu: one resident wave of networks at 2N=402."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from tc_gan_b200 import torch_ops as ops, ssnode, stimuli
dev = torch.device('cuda:0')
P = ssnode.DEFAULT_PARAMS; jds = ssnode.new_JDS()
n_sites = 201; dim = 402
nz = int(os.environ.get('NZ', 37)); seqlen = int(os.environ.get('SEQLEN', 300)); skip = seqlen - 100
exts = torch.tensor(stimuli.input(P['bandwidths'], np.linspace(-.5, .5, n_sites), P['smoothness'], P['contrast']), dtype=torch.float32, device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(0)
z = torch.rand((nz, dim, dim), generator=gen, device=dev)
G = torch.randn((nz, 8, dim), generator=gen, device=dev)
J, D, S = (torch.tensor(np.asarray(jds[k]), dtype=torch.float64, device=dev, requires_grad=True) for k in 'JDS')
for rep in range(int(os.environ.get('REPS', 2))):
    for p in (J, D, S): p.grad = None
    R, st, it = ops.ssn_fixed_point(z, J, D, S, exts)
    (R * G).sum().backward()
    avg, dyn, rate = ops.euler_ssn(z, J, D, S, exts, seqlen=seqlen, skip_steps=skip)
    ((avg * G).sum() + 0.1 * dyn + 0.01 * rate).backward()
torch.cuda.synchronize()
print('ok', int((st == 0).sum()), float(J.grad.abs().sum()))
if os.environ.get('F64'):
    # the reference-ABI symbol (float64 cluster kernel): one (network, stimulus) per call, and one batched precise call
    from tc_gan_b200 import weight_gen
    zz = z[0].double().cpu().numpy()
    W = weight_gen.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], zz)
    e64 = exts.double().cpu().numpy()
    sol = ssnode.fixed_point(W, e64[7], k=P['k'], n=P['n'])
    Rs, errs, its = ssnode.fixed_points_batch(np.stack([W] * 8), e64, precise=True, k=P['k'], n=P['n'])
    print('f64 ok', sol.error, int((errs == 0).sum()))
