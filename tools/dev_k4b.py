"""K4b check: tcgen05 kernel vs the FFMA kernel vs float64 on random data (development)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import numpy as np, torch
import ssn_oracle as so
from tc_gan_b200 import clib, ssnode, torch_ops as ops
dev = torch.device('cuda:0')
def run(n_sites, nz, nb, seqlen, reps=1):
    dim = 2 * n_sites; pitch = clib.libssnode.ssn_traj_pitch(n_sites)
    jds = ssnode.new_JDS()
    g = torch.Generator(device=dev); g.manual_seed(n_sites + nz)
    z = torch.rand((nz, dim, dim), generator=g, device=dev)
    traj = torch.full((nz, seqlen, nb, pitch), float('nan'), device=dev)
    adj = torch.full((nz, seqlen, nb, pitch), float('nan'), device=dev)
    traj[..., :dim] = torch.rand((nz, seqlen, nb, dim), generator=g, device=dev) * 10
    adj[..., :dim] = torch.randn((nz, seqlen, nb, dim), generator=g, device=dev)
    # float64 reference of the contraction + theta reduction
    A = adj[..., :dim].double().reshape(nz, -1, dim); B = traj[..., :dim].double().reshape(nz, -1, dim)
    G = torch.einsum('zki,zkj->zij', A, B).cpu().numpy()
    dJ, dD, dS = so.weight_param_contraction(n_sites, jds['J'], jds['D'], jds['S'], z.double().cpu().numpy(), G)
    want = np.concatenate([dJ.ravel(), dD.ravel(), dS.ravel()])
    if os.environ.get('SSN_K4B_DEBUG'):
        a0 = adj[0].reshape(-1, pitch).cpu().numpy(); b0 = traj[0].reshape(-1, pitch).cpu().numpy()
        print('expect A row0', a0[0, :8], '\n       A row1', a0[1, :8], '\n       B row0', b0[0, :8])
        print('expect acc row0', G[0, 0, :8], '\n       acc row1', G[0, 1, :8], '\n       acc row5 c16', G[0, 5, 16:24])
    # call the backward entry with a zero adjoint recursion?  no: call the contraction through a tiny shim
    out = {}
    for mode in ('tc', 'ffma'):
        os.environ['SSN_K4B'] = mode
        grad = torch.zeros(12, dtype=torch.float64, device=dev)
        rc = clib.libssnode.ssn_bptt_param_grad(nz, nb, n_sites, seqlen, adj.data_ptr(), traj.data_ptr(), z.data_ptr(),
                                               clib.make_jds(jds['J'], jds['D'], jds['S']), grad.data_ptr(), None)
        clib.check_call(rc, 'ssn_bptt_param_grad')
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            clib.libssnode.ssn_bptt_param_grad(nz, nb, n_sites, seqlen, adj.data_ptr(), traj.data_ptr(), z.data_ptr(),
                                               clib.make_jds(jds['J'], jds['D'], jds['S']), grad.data_ptr(), None)
        e1.record(); torch.cuda.synchronize()
        got = grad.cpu().numpy()
        err = np.abs(got - want).max() / np.abs(want).max()
        out[mode] = (err, e0.elapsed_time(e1) / reps)
    print('n_sites %d nz %d nb %d seqlen %d: tc err %.2e %.3f ms | ffma err %.2e %.3f ms' % (
        n_sites, nz, nb, seqlen, out['tc'][0], out['tc'][1], out['ffma'][0], out['ffma'][1]), flush=True)
for shape in ((10, 1, 3, 5), (51, 2, 8, 30), (201, 1, 8, 12), (201, 3, 8, 200)):
    run(*shape)
if os.environ.get('FULL'):
    run(201, 128, 8, 1200, reps=3)
