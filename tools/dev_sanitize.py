"""Tiny run of every kernel for compute-sanitizer memcheck."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from tc_gan_b200 import torch_ops as ops, ssnode, stimuli
dev = torch.device('cuda:0')
jds = ssnode.new_JDS(); P = ssnode.DEFAULT_PARAMS
for n_sites, nz, nb in ((40, 3, 9), (201, 2, 8)):
    dim = 2 * n_sites
    exts = torch.tensor(stimuli.input(np.linspace(0.1, 1, nb), np.linspace(-.5, .5, n_sites), P['smoothness'], P['contrast']), dtype=torch.float32, device=dev)
    z = torch.rand((nz, dim, dim), device=dev)
    J, D, S = (torch.tensor(jds[k], dtype=torch.float64, device=dev, requires_grad=True) for k in 'JDS')
    R, st, it = ops.ssn_fixed_point(z, J, D, S, exts, solver=ops.make_solver(max_iter=400))
    (R * torch.randn_like(R)).sum().backward()
    os.environ['SSN_FORCE_SMEM_KERNEL'] = '1'
    ops.fixed_points(z, J, D, S, exts, solver=ops.make_solver(max_iter=100))
    del os.environ['SSN_FORCE_SMEM_KERNEL']
    ops.fixed_points(z, J, D, S, exts, solver=ops.make_solver(max_iter=60), precise=True)
    J.grad = None
    avg, dyn, rate = ops.euler_ssn(z, J, D, S, exts, seqlen=12, skip_steps=6)
    (avg.sum() + dyn + rate).backward()
    torch.cuda.synchronize()
    print('ok', n_sites, int((st == 0).sum()))
