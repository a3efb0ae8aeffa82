"""K2 development: accuracy (vs the float64 oracle) and time of the implicit-gradient kernel; SSN_IFT=damped for the
round-1 damped iteration."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import numpy as np, torch
import ssn_oracle as so
from tc_gan_b200 import torch_ops as ops, ssnode, stimuli
dev = torch.device('cuda:0')
P = ssnode.DEFAULT_PARAMS; jds = ssnode.new_JDS()
def t(a, dt=torch.float32, g=False):
    x = torch.tensor(np.asarray(a), dtype=dt, device=dev); return x.requires_grad_() if g else x
for n_sites, nz, nchk in ((201, int(os.environ.get('NZ', 256)), 2), (51, 64, 4), (20, 16, 4)):
    dim = 2 * n_sites
    exts = stimuli.input(P['bandwidths'], np.linspace(-.5, .5, n_sites), P['smoothness'], P['contrast'])
    gen = torch.Generator(device=dev); gen.manual_seed(0)
    z = torch.rand((nz, dim, dim), generator=gen, device=dev)
    J, D, S = (t(jds[k], torch.float64) for k in 'JDS')
    G = torch.randn((nz, 8, dim), generator=gen, device=dev)
    R, st, it = ops.fixed_points(z, J, D, S, t(exts))
    torch.cuda.synchronize()
    def run():
        return ops.ift_gradient(z, J, D, S, t(exts), R, G, return_mu=True)
    out = run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = run()
    e1.record(); torch.cuda.synchronize()
    gJ, gD, gS, mu, status, iters = out
    print('2N=%d nz=%d: %.2f ms  status!=0: %d  sweeps mean %.1f max %d' % (dim, nz, e0.elapsed_time(e1) / 3, int((status != 0).sum()), iters.float().mean(), iters.max()))
    zs = z[:nchk].double().cpu().numpy(); Rs = R[:nchk].double().cpu().numpy(); Gs = G[:nchk].double().cpu().numpy()
    W = so.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], zs)
    dJ, dD, dS, mu_o = so.ift_param_gradient(Rs, W, zs, exts, jds['J'], jds['D'], jds['S'], Gs)
    g2 = ops.ift_gradient(z[:nchk], J, D, S, t(exts), R[:nchk], G[:nchk], return_mu=True)
    print('   mu max err / max|mu| = %.2e' % (np.abs(g2[3].cpu().numpy() - mu_o).max() / np.abs(mu_o).max()))
    for name, got, want in (('J', g2[0], dJ), ('D', g2[1], dD), ('S', g2[2], dS)):
        print('   d%s max rel err %.2e' % (name, np.abs(got.cpu().numpy() - want).max() / np.abs(want).max()))
