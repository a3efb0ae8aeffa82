"""Timing + accuracy of the gradient paths at BASELINE configs 3 and 4 (development)."""
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
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
n_sites = 201; dim = 402
exts = stimuli.input(P['bandwidths'], np.linspace(-.5, .5, n_sites), P['smoothness'], P['contrast'])
gen = torch.Generator(device=dev); gen.manual_seed(0)
# ---- config 4: fixed-point generator step, 256 networks x 8 stimuli ----
nz = int(os.environ.get('NZ4', 256))
z = torch.rand((nz, dim, dim), generator=gen, device=dev)
J, D, S = (t(jds[k], torch.float64, True) for k in 'JDS')
G = torch.randn((nz, 8, dim), generator=gen, device=dev)
def fp_step():
    for p in (J, D, S): p.grad = None
    R, st, it = ops.ssn_fixed_point(z, J, D, S, t(exts))
    (R * G).sum().backward()
    return R, st
ms = timed(fp_step)
R, st = fp_step()
print('config 4: fixed-point fwd + IFT bwd, %d networks x 8: %.1f ms/step -> %.2f gen steps/s (converged %d/%d)' % (nz, ms, 1e3 / ms, int((st == 0).sum()), st.numel()))
ms_f = timed(lambda: ops.fixed_points(z, J, D, S, t(exts)))
print('   forward only %.1f ms' % ms_f)
# accuracy of the IFT gradient on 2 networks vs the float64 oracle
zs = z[:2].double().cpu().numpy(); Rs = R[:2].detach().double().cpu().numpy(); Gs = G[:2].double().cpu().numpy()
W = so.generate_weight(n_sites, jds['J'], jds['D'], jds['S'], zs)
dJ, dD, dS, mu = so.ift_param_gradient(Rs, W, zs, exts, jds['J'], jds['D'], jds['S'], Gs)
gJ, gD, gS = ops.ift_gradient(z[:2], J, D, S, t(exts), R[:2].detach(), G[:2])
for name, got, want in (('J', gJ, dJ), ('D', gD, dD), ('S', gS, dS)):
    print('   IFT d%s max rel err %.2e' % (name, np.abs(got.cpu().numpy() - want).max() / np.abs(want).max()))
# ---- config 3: BPTT generator step, 128 networks x 8 stimuli, seqlen 1200 ----
nz3 = int(os.environ.get('NZ3', 128)); seqlen = int(os.environ.get('SEQLEN', 1200)); skip = seqlen - 200
z3 = torch.rand((nz3, dim, dim), generator=gen, device=dev)
G3 = torch.randn((nz3, 8, dim), generator=gen, device=dev)
def bptt_step():
    for p in (J, D, S): p.grad = None
    avg, dyn, rate = ops.euler_ssn(z3, J, D, S, t(exts), seqlen=seqlen, skip_steps=skip)
    ((avg * G3).sum() + 0.1 * dyn + 0.01 * rate).backward()
ms = timed(bptt_step, reps=2)
print('config 3: BPTT fwd+bwd, %d networks x 8, seqlen %d: %.1f ms/step -> %.2f gen steps/s' % (nz3, seqlen, ms, 1e3 / ms))
ms_f = timed(lambda: ops.euler_forward(z3, J, D, S, t(exts), seqlen, skip, ops.clib.make_solver(tau=(10., 1.), dt=0.1), store=True), reps=2)
print('   forward only %.1f ms' % ms_f)
