"""
bench.py -- converged SSN solves/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py --gpus N --steps K --warmup W                     # our CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W    # reference C solver on host cores

Workload (BASELINE.json configs[1]): ring SSN with n_sites=201 (2N=402), 1024 sampled networks x 8 stimuli per
GPU and per step; new_JDS parameters, asym_tanh, k=0.01, n=2.2, tau=(0.01589, 0.002), dt=8e-4, atol=1e-5, r0=0
(SURVEY.md section 8d).  z ~ U[0,1) is synthetic: network i of rank r is
``np.random.RandomState(SEED + r).rand(2N, 2N)`` drawn i-th from that stream, rounded to float32 -- the SAME list
in both arms (the reference arm and the cpu_baseline leg solve a prefix of rank 0's list), and the GPU results
of that prefix are checked against the reference solver's in the run (`parity`).  W is built on chip from z.

* value : converged (status 0) solves of all ranks / max-over-ranks device time, inputs resident in HBM
          (CUDA events on the launching stream).
* e2e   : the same metric through the C ABI with HOST buffers: pinned float32 z in, R/status out,
          host<->device copies inside the timed region (`ssn_fixed_point_batch`, SSN_MEM_HOST).
          `e2e.api` is the same through the reference-facing Python API, `ssnode.find_fixed_points(num,
          Z_W_gen, exts)` with one float64 W per network exactly as tc_gan/run/gan.py:622-634 calls it, and
          `e2e.api_z` its z-based variant (jds=..., float32 z, W built on the GPU).
* roofline : FP32 FFMA.  achieved = sum over solves of sweeps x 2 (2N)^2 flop / average duration of the
          fixed-point kernel (CUDA events around that launch, `ssn_profile_enable`); peak = nominal
          148 SM x 128 FMA x 2 x max SM clock (MEASURED_PEAKS.json carries only HBM and bf16 numbers; the
          FFMA throughput measured by the library's probe kernel is reported beside it).  traffic = DRAM bytes
          of the kernel per launch from the committed ncu capture named in profiles/k1_dram_traffic.json.
* cpu_baseline / --impl reference : the UNMODIFIED reference C solver (oracle/_ref/libssnode.so) on every host
          core, driven by a thread pool as tc_gan.ssnode.find_fixed_points_parallel drives it
          (oracle/ssn_oracle.py:ref_fixed_point_batch), on a bounded prefix of the same network list.
* secondary : GAN generator steps/s (configs[2] BPTT and configs[3] fixed-point, the step's networks sharded
          over the ranks, ONE NCCL all-reduce of the packed (dJ, dD, dS) inside the timed loop), per-kernel
          device time and FP32-roofline fraction of K2 / K3 / K4 / K4b, and a 50-stimulus slice (configs[4]
          shape).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SITES, NZ, NB_BANDWIDTHS = 201, 1024, [0, 0.0625, 0.125, 0.1875, 0.25, 0.5, 0.75, 1]
METRIC = 'converged SSN solves/sec (2N=402, 8 stim)'
UNIT = 'solves/s'
SEED = 20251018
DIM = 2 * N_SITES
FLOP_PER_SWEEP = 2.0 * DIM * DIM


def workload_config(nz):
    """Identical in both arms (the driver compares the dicts)."""
    return {'workload': 'configs[1]: ssnode fixed-point solve, n_sites=201 (2N=402), %d networks x 8 stimuli per GPU per step' % nz,
            'io_type': 'asym_tanh', 'params': 'new_JDS', 'dt': 8e-4, 'atol': 1e-5, 'max_iter': 10000,
            'networks_per_gpu': nz, 'stimuli': 8, 'z_seed': SEED,
            'l2': 'inputs larger than L2 (z = %.0f MB per step)' % (nz * DIM * DIM * 4 / 1e6)}


def draw_networks(n, rank=0):
    """The benchmark's network list: float32 z [n, 2N, 2N]; a shorter list is a prefix of a longer one."""
    import numpy as np
    rs = np.random.RandomState(SEED + rank)
    out = np.empty((n, DIM, DIM), dtype=np.float32)
    for i in range(n):
        out[i] = rs.rand(DIM, DIM)
    return out


class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons),
                'samples': len(sm)}


def host_cores():
    """Threads for the CPU arms: every core this process may run on.  (tc_gan.utils.cpu_count honours
    OMP_NUM_THREADS, but torchrun exports OMP_NUM_THREADS=1 to its ranks, which would cripple the
    reference arm; SSN_BASELINE_THREADS overrides.)"""
    if os.environ.get('SSN_BASELINE_THREADS'):
        return max(1, int(os.environ['SSN_BASELINE_THREADS']))
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def reference_solver():
    """(oracle module, kind, callable(W, exts, threads) -> (R, status)) for the CPU arms."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ssn_oracle as so
    if so.ref_lib() is not None:
        return so, 'reference', lambda W, exts, threads: so.ref_fixed_point_batch(W, exts, threads=threads)

    def port(W, exts, threads):
        R, st, _ = so.fixed_point_batch(W, exts, threads=threads, stop_at_first_failure=True)
        return R, st
    return so, 'port', port


CPU_DRIVER = ('thread pool over the unmodified reference C symbol solve_dynamics_asym_tanh_euler, one network per '
              'job, stimuli last to first, as tc_gan.ssnode.find_fixed_points_parallel drives it '
              '(oracle/ssn_oracle.py:ref_fixed_point_batch; tc_gan.ssnode itself needs Theano to import)')


def run_reference(args):
    """The reference's own CPU implementation on the host cores (rank 0 only)."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    import numpy as np
    so, kind, solve = reference_solver()
    cores = host_cores()
    nz = max(8, min(NZ, 32 * cores))            # bounded prefix of the 1024-network step (~5 s of CPU work per step)
    jds = so.new_JDS()
    exts = so.stimulus_input(so.DEFAULT_BANDWIDTHS, N_SITES)
    W = so.generate_weight(N_SITES, jds['J'], jds['D'], jds['S'], draw_networks(nz).astype(np.float64))

    def step():
        _, st = solve(W, exts, cores)
        return int((st == 0).sum())

    for _ in range(args.warmup):
        step()
    t0 = time.time()
    solved = sum(step() for _ in range(args.steps))
    dt = time.time() - t0
    value = solved / dt
    sample = 'first %d of the %d networks x 8 stimuli per step, %d steps' % (nz, NZ, args.steps)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(NZ),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample,
                         'driver': CPU_DRIVER},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


def cpu_baseline_and_parity(R_gpu, status_gpu):
    """Times the reference solver on a prefix of rank 0's network list and checks the GPU results of that prefix
    against it: equal status codes, rates within rtol 1e-4 (BASELINE north_star)."""
    import numpy as np
    so, kind, solve = reference_solver()
    cores = host_cores()
    nz = max(8, min(NZ, 64 * cores, len(R_gpu)))             # ~10 s of CPU work on all host cores
    jds = so.new_JDS()
    exts = so.stimulus_input(so.DEFAULT_BANDWIDTHS, N_SITES)
    W = so.generate_weight(N_SITES, jds['J'], jds['D'], jds['S'], draw_networks(nz).astype(np.float64))
    t0 = time.time()
    R_ref, st = solve(W, exts, cores)
    dt = time.time() - t0
    base = {'value': float((st == 0).sum() / dt), 'unit': UNIT, 'cores': cores, 'kind': kind,
            'sample': 'first %d of the %d networks x 8 stimuli, one pass (%.1f s)' % (nz, NZ, dt),
            'driver': CPU_DRIVER}
    ok = st == 0
    err = np.abs(R_gpu[:nz] - R_ref)[ok] / (1e-4 * np.abs(R_ref)[ok] + 1e-4)
    parity = {'networks': nz, 'solves': int(st.size), 'status_equal': bool((status_gpu[:nz] == st).all()),
              'max_err_over_tol': float(err.max()) if err.size else 0.0, 'rtol': 1e-4, 'atol': 1e-4,
              'against': kind}
    parity['ok'] = parity['status_equal'] and parity['max_err_over_tol'] <= 1.0
    return base, parity


# ---------------------------------------------------------------------------------------------------
# secondary: GAN generator steps (configs[2], configs[3]) and the 50-stimulus slice (configs[4] shape)
# ---------------------------------------------------------------------------------------------------

def dist_env():
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')),
            int(os.environ.get('WORLD_SIZE', '1')))


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        return {}


def fp32_peak_tflops(sm_max_mhz):
    return 148 * 128 * 2 * (sm_max_mhz or 1965.0) * 1e6 * 1e-12


def time_gan_step(workload, steps, warmup, dev, world, rank, peak):
    """One generator step = forward + backward through the SSN with a fixed linear critic, the step's networks
    sharded over the ranks (strong scaling), one all-reduce of the packed 12-double gradient per step.
    Returns the dict for the JSON line (rank 0) -- device time by CUDA events, max over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from tc_gan_b200 import clib, ssnode, stimuli, torch_ops as ops
    from tc_gan_b200 import dist as sdist
    P = ssnode.DEFAULT_PARAMS
    jds = ssnode.new_JDS()
    exts = torch.tensor(stimuli.input(NB_BANDWIDTHS, np.linspace(-.5, .5, N_SITES), P['smoothness'], P['contrast']),
                        dtype=torch.float32, device=dev)
    total = 256 if workload == 'gan_fp' else 128
    seqlen, skip = 1200, 1000
    nz = len(sdist.shard_indices(total, rank, world))
    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + rank)
    z = torch.rand((nz, DIM, DIM), generator=gen, device=dev)
    G = torch.randn((nz, exts.shape[0], DIM), generator=gen, device=dev)
    J, D, S = (torch.tensor(jds[k], dtype=torch.float64, device=dev, requires_grad=True) for k in 'JDS')
    stats = {}

    def step():
        for p in (J, D, S):
            p.grad = None
        if workload == 'gan_fp':
            R, status, iters = ops.ssn_fixed_point(z, J, D, S, exts)
            loss = (R * G).sum()
            stats['iters'] = iters
        else:
            avg, dyn, rate = ops.euler_ssn(z, J, D, S, exts, seqlen=seqlen, skip_steps=skip)
            loss = (avg * G).sum() + 0.1 * dyn + 0.01 * rate
        loss.backward()
        return sdist.allreduce_generator_grads(J.grad, D.grad, S.grad)

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clib.profile_enable(True)
    clib.profile_read()
    launches0 = clib.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    kernels = clib.profile_read()
    clib.profile_enable(False)
    t_end = time.time()
    launches = clib.kernel_launches() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item() / steps
    # algorithmic flop of rank 0's kernels (per launch) for their FP32-roofline fractions
    per_kernel = {}
    flop = {}
    if workload == 'gan_fp':
        sweeps = float(stats['iters'].to(torch.int64).sum().item())
        flop['ssn_fp_ws_kernel'] = sweeps * FLOP_PER_SWEEP
        # K2: per (network, stimulus) one contraction for Phi + the adjoint sweeps it reports; per network the
        # dL/dW outer products of the 8 stimuli and the 12-way reduction (~26 (2N)^2 flop)
        adj_iters = ops.last_adjoint['iters']
        adj_sweeps = float(adj_iters.to(torch.int64).sum().item()) if adj_iters is not None else 0.0
        flop['ssn_ift_cluster_kernel'] = (adj_sweeps + nz * 8 + nz * 13) * FLOP_PER_SWEEP
        stats['adj_sweeps_per_solve'] = adj_sweeps / max(nz * 8, 1)
    else:
        flop['ssn_euler_cluster_kernel_fwd'] = nz * 8 * seqlen * FLOP_PER_SWEEP
        flop['ssn_euler_cluster_kernel_bwd'] = nz * 8 * (seqlen - 1) * FLOP_PER_SWEEP
        flop['ssn_bptt_param_grad_kernel'] = nz * 8 * seqlen * FLOP_PER_SWEEP
        flop['ssn_bptt_param_grad_tc_kernel'] = nz * 8 * seqlen * FLOP_PER_SWEEP
    for name, (tot_ms, n) in sorted(kernels.items()):
        ent = {'ms_per_launch': tot_ms / n, 'launches_per_step': n / steps}
        if flop.get(name):
            ent['tflops'] = flop[name] / (tot_ms / n * 1e-3) * 1e-12
            ent['fp32_roofline_frac'] = ent['tflops'] / peak
        if name.startswith('ssn_euler_cluster_kernel') and flop.get(name):
            # shared-memory view (north_star: 'fraction of the FP32/SMEM roofline'): per time step a CTA of the 4-CTA
            # cluster issues, per warp and column step, 7 LDS.128 of W and 8 of the panel = 15 x 512 B, 7 column steps,
            # 8 warps; peak 128 B per clock and SM
            steps_done = seqlen if name.endswith('fwd') else seqlen - 1
            smem_bytes = 15 * 512.0 * 7 * 8 * 4 * steps_done * nz
            smem_peak = 148 * 128.0 * (peak / (148 * 128 * 2.0)) * 1e12          # B/s at the clock of the FP32 peak
            ent['smem'] = {'achieved_gbs': smem_bytes / (tot_ms / n * 1e-3) * 1e-9, 'peak_gbs': smem_peak * 1e-9,
                           'frac': smem_bytes / (tot_ms / n * 1e-3) / smem_peak}
        if name == 'ssn_bptt_param_grad_tc_kernel' and flop.get(name):
            # tensor-pipe view: three kind::tf32 MMAs per product (hi/lo split) on 128 x 208 tiles of the 402 x 402
            # output (512 x 416 computed); peak = half the measured dense bf16 rate (TF32 runs at half the bf16 rate;
            # nominal 1.1 PFLOP/s, /opt/skills/guides/B200_PROFILING.md)
            executed = ent['tflops'] * 3.0 * (512.0 * 416.0) / float(DIM * DIM)
            tpeak = measured_peaks().get('bf16_tflops_sustained', 2250.0) / 2.0
            ent['tensor'] = {'executed_tflops': executed, 'peak': tpeak, 'unit': 'TFLOP/s', 'frac': executed / tpeak,
                             'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained / 2'}
        per_kernel[name] = ent
    return {'metric': 'GAN generator steps/sec (%s)' % ('fixed-point, implicit gradient' if workload == 'gan_fp'
                                                        else 'BPTT, seqlen 1200'),
            'value': 1e3 / ms_step, 'unit': 'steps/s', 'ms_per_step': ms_step, 'steps': steps, 'scaling': 'strong',
            'workload': ('configs[3]: fixed-point GAN generator step, 256 networks x 8 stimuli, 2N=402' if workload == 'gan_fp'
                         else 'configs[2]: bptt_cwgan generator step, 128 networks x 8 stimuli, 2N=402, seqlen 1200, skip 1000'),
            'networks_per_gpu': nz, 'critic': 'fixed linear functional',
            'collective': 'one all-reduce of the packed 12-double (dJ, dD, dS) per step, inside the timed loop',
            'gpu_launches': launches, 'kernels': per_kernel,
            **({'adjoint_sweeps_per_solve': stats['adj_sweeps_per_solve']} if 'adj_sweeps_per_solve' in stats else {})}


def time_nb50_slice(dev, peak, nz=128):
    """configs[4] shape on one GPU: `nz` networks x 50 stimuli (5 contrasts x 10 bandwidths), device-resident."""
    import numpy as np
    import torch
    from tc_gan_b200 import clib, ssnode, stimuli
    P = ssnode.DEFAULT_PARAMS
    jds = ssnode.new_JDS()
    exts_np = stimuli.input(np.linspace(0, 1, 10), np.linspace(-.5, .5, N_SITES), P['smoothness'], [5, 10, 20, 30, 40])
    nb = len(exts_np)
    gen = torch.Generator(device=dev)
    gen.manual_seed(4242)
    z = torch.rand((nz, DIM, DIM), generator=gen, device=dev)            # device-side Philox z
    ext = torch.tensor(exts_np, dtype=torch.float32, device=dev)
    R = torch.empty((nz, nb, DIM), dtype=torch.float32, device=dev)
    status = torch.empty((nz, nb), dtype=torch.int32, device=dev)
    iters = torch.empty((nz, nb), dtype=torch.int32, device=dev)
    sv = clib.make_solver(k=P['k'], n=P['n'])
    jd = clib.make_jds(jds['J'], jds['D'], jds['S'])
    stream = torch.cuda.current_stream()

    def step():
        clib.check_call(clib.libssnode.ssn_fixed_point_batch(
            sv, nz, nb, N_SITES, clib.W_FROM_Z, z.data_ptr(), jd, ext.data_ptr(), 0, None, R.data_ptr(),
            status.data_ptr(), iters.data_ptr(), 0, clib.MEM_DEVICE, stream.cuda_stream), 'ssn_fixed_point_batch')

    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    conv = int((status == 0).sum().item())
    sweeps = float(iters.to(torch.int64).sum().item())
    tf = sweeps * FLOP_PER_SWEEP / (ms * 1e-3) * 1e-12
    return {'workload': 'configs[4] shape: %d networks x 50 stimuli (5 contrasts x 10 bandwidths), 2N=402, device-side z' % nz,
            'value': conv / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'converged': conv, 'solves': nz * nb,
            'mean_sweeps_per_solve': sweeps / (nz * nb), 'fp32_roofline_frac': tf / peak}


def run_sweep(args):
    """`--workload sweep`: configs[4] at full size -- 65536 networks x 50 stimuli at 2N=402, z drawn on the device
    (torch Philox, a fresh slab per launch), the networks split evenly over the ranks (strong scaling, no collective
    on the data path; the converged counts and the max time are reduced once after the timed region)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from tc_gan_b200 import clib, ssnode, stimuli
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    total, slab = args.sweep_networks, 512
    mine = total // world + (1 if rank < total % world else 0)
    P = ssnode.DEFAULT_PARAMS
    jds = ssnode.new_JDS()
    exts_np = stimuli.input(np.linspace(0, 1, 10), np.linspace(-.5, .5, N_SITES), P['smoothness'], [5, 10, 20, 30, 40])
    nb = len(exts_np)
    ext = torch.tensor(exts_np, dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(777 + rank)
    z = [torch.empty((slab, DIM, DIM), device=dev) for _ in range(2)]
    R = torch.empty((slab, nb, DIM), dtype=torch.float32, device=dev)
    status = torch.empty((slab, nb), dtype=torch.int32, device=dev)
    iters = torch.empty((slab, nb), dtype=torch.int32, device=dev)
    conv = torch.zeros((), dtype=torch.int64, device=dev)
    sweeps = torch.zeros((), dtype=torch.int64, device=dev)
    sv = clib.make_solver(k=P['k'], n=P['n'])
    jd = clib.make_jds(jds['J'], jds['D'], jds['S'])
    stream = torch.cuda.current_stream()

    def solve(zz, n):
        clib.check_call(clib.libssnode.ssn_fixed_point_batch(
            sv, n, nb, N_SITES, clib.W_FROM_Z, zz.data_ptr(), jd, ext.data_ptr(), 0, None, R.data_ptr(),
            status.data_ptr(), iters.data_ptr(), 0, clib.MEM_DEVICE, stream.cuda_stream), 'ssn_fixed_point_batch')

    z[0].uniform_(generator=gen)
    solve(z[0], min(slab, 64))                                  # warm-up (module load, occupancy queries)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local_rank)
    t_start = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = clib.kernel_launches()
    e0.record()
    done, i = 0, 0
    while done < mine:
        n = min(slab, mine - done)
        z[i & 1].uniform_(generator=gen)                        # Philox z on the device, inside the timed region
        solve(z[i & 1], n)
        conv += (status[:n] == 0).sum()
        sweeps += iters[:n].sum(dtype=torch.int64)
        done += n
        i += 1
    e1.record()
    torch.cuda.synchronize()
    t_end = time.time()
    launches = clib.kernel_launches() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    tot = torch.stack([conv, sweeps]).to(torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    clk = clocks.stop(t_start, t_end)
    if rank == 0:
        sec = ms.item() * 1e-3
        peak = fp32_peak_tflops(clk.get('sm_max_mhz')) * world
        tf = tot[1].item() * FLOP_PER_SWEEP / sec * 1e-12
        print(json.dumps({
            'metric': 'converged SSN solves/sec (2N=402, 50 stim)', 'value': tot[0].item() / sec, 'unit': UNIT,
            'n_gpus': world, 'steps': 1, 'warmup': 1, 'ms_per_step': ms.item(), 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32 contraction / f64 state', 'data': 'synthetic',
            'config': {'workload': 'configs[4]: throughput sweep, %d networks x %d stimuli (5 contrasts x 10 bandwidths), '
                                   '2N=402, device-side Philox z, slabs of %d networks' % (total, nb, slab),
                       'networks_per_gpu': mine, 'l2': 'a fresh z slab (331 MB) per launch'},
            'converged': int(tot[0].item()), 'solves': total * nb, 'mean_sweeps_per_solve': tot[1].item() / (total * nb),
            'gpu_launches': launches, 'clocks': clk,
            'roofline': {'bound': 'fp32_ffma', 'achieved': tf, 'peak': peak, 'unit': 'TFLOP/s', 'frac': tf / peak}}))
    if world > 1:
        dist.destroy_process_group()


def run_secondary_only(args):
    """`--workload gan_fp|gan_bptt`: the secondary metric as the line's own value."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    res = time_gan_step(args.workload, args.steps, max(args.warmup, 3), dev, world, rank, fp32_peak_tflops(None))
    if rank == 0:
        res.update({'n_gpus': world, 'warmup': max(args.warmup, 3), 'higher_is_better': True, 'vs_baseline': None,
                    'dtype': 'f32 contraction / f64 state', 'data': 'synthetic', 'config': {'workload': res.pop('workload')}})
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


def run_wgan(args):
    """Full WGAN-GP training steps (5 critic updates + 1 generator update per step) with the torch MLP
    critic of tc_gan_b200.gan around the CUDA generator; 'true' tuning curves sampled from the true-parameter SSN."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from tc_gan_b200 import clib, gan
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    mode = 'fixed_point' if args.workload == 'wgan_fp' else 'bptt'
    num_models = 256 if mode == 'fixed_point' else 128
    # 'true' tuning curves from the true-parameter SSN through the solver (as run/gan.py draws its data set), the
    # generator starting from perturbed parameters (init_disturbance of run/gan.py:345-352)
    from tc_gan_b200 import ssnode
    jds = ssnode.new_JDS()
    data, _ = ssnode.sample_tuning_curves(sample_sites=[0], NZ=512, seed=1, N=N_SITES, J=jds['J'], D=jds['D'], S=jds['S'])
    data = np.asarray(data, dtype=np.float32)
    if data.shape[0] == 8 and data.shape[1] != 8:
        data = data.T
    g = gan.SSNWassersteinGAN(data, num_sites=N_SITES, mode=mode, num_models=num_models, sample_sites=(0,),
                              J=jds['J'] * 0.9, D=jds['D'] * 0.9, S=jds['S'] * 0.95,
                              critic_iters_init=5, critic_iters=5, device=dev)
    steps = g.learning()

    def gen_step():
        for info in steps:
            if not info['is_discriminator']:
                return info

    for _ in range(max(1, args.warmup // 3)):
        gen_step()
    torch.cuda.synchronize()
    launches0 = clib.kernel_launches()
    t0 = time.time()
    for _ in range(args.steps):
        info = gen_step()
    torch.cuda.synchronize()
    dt = torch.tensor([time.time() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({
            'metric': 'WGAN-GP training steps/sec (5 critic + 1 generator update, %s generator)' % mode,
            'value': args.steps / dt.item(), 'unit': 'steps/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt.item() / args.steps * 1e3, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32 contraction / f64 state', 'data': 'synthetic',
            'config': {'workload': '%d networks x 8 stimuli per update, 2N=402, critic MLP 128-128' % num_models,
                       'last_gen_loss': info['gen_loss'], 'rejections': g.rejections},
            'gpu_launches': clib.kernel_launches() - launches0}))
    if world > 1:
        dist.destroy_process_group()


def k1_dram_traffic(kernel_tag):
    """DRAM bytes per network of the fixed-point kernel from the committed ncu capture
    (profiles/k1_dram_traffic.json, keyed by the kernel's shape tag); None when no capture matches."""
    try:
        table = json.load(open(os.path.join(ROOT, 'profiles', 'k1_dram_traffic.json')))
    except (OSError, ValueError):
        return None, None
    ent = table.get(kernel_tag)
    if not ent:
        return None, None
    return float(ent['dram_bytes_per_network']), ent.get('capture')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--networks', type=int, default=NZ, help='networks per GPU per step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip the GAN-step / 50-stimulus secondary block')
    ap.add_argument('--workload', default='solve', choices=['solve', 'gan_fp', 'gan_bptt', 'wgan_fp', 'wgan_bptt', 'sweep'],
                    help='solve: configs[1] (default, the BASELINE metric, with the secondary block); gan_fp / gan_bptt: '
                         'only that generator step; wgan_*: full WGAN-GP training steps; sweep: configs[4], the '
                         '65536-network x 50-stimulus throughput sweep split over the ranks')
    ap.add_argument('--sweep-networks', type=int, default=65536, help='total networks of --workload sweep')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    if args.workload == 'sweep':
        return run_sweep(args)
    if args.workload in ('wgan_fp', 'wgan_bptt'):
        return run_wgan(args)
    if args.workload != 'solve':
        return run_secondary_only(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from tc_gan_b200 import clib, ssnode, stimuli

    rank, local_rank, world = dist_env()
    if clib.libssnode.ssn_device_count() < 1:
        raise SystemExit('bench.py: no CUDA device; the SSN library has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    nz, n_sites = args.networks, N_SITES
    P = ssnode.DEFAULT_PARAMS
    jds = ssnode.new_JDS()
    exts_np = stimuli.input(NB_BANDWIDTHS, np.linspace(-.5, .5, n_sites), P['smoothness'], P['contrast'])
    nb = len(exts_np)
    sv = clib.make_solver(k=P['k'], n=P['n'])
    jd = clib.make_jds(jds['J'], jds['D'], jds['S'])
    lib = clib.libssnode

    z_np = draw_networks(nz, rank)                               # every rank owns different networks
    z_host = torch.empty((nz, DIM, DIM), dtype=torch.float32, pin_memory=True)
    z_host.copy_(torch.from_numpy(z_np))
    z = z_host.to(dev)
    ext = torch.tensor(exts_np, dtype=torch.float32, device=dev)
    R = torch.empty((nz, nb, DIM), dtype=torch.float32, device=dev)
    status = torch.empty((nz, nb), dtype=torch.int32, device=dev)
    iters = torch.empty((nz, nb), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def step_device():
        clib.check_call(lib.ssn_fixed_point_batch(
            sv, nz, nb, n_sites, clib.W_FROM_Z, z.data_ptr(), jd, ext.data_ptr(), 0, None,
            R.data_ptr(), status.data_ptr(), iters.data_ptr(), 0, clib.MEM_DEVICE, stream.cuda_stream),
            'ssn_fixed_point_batch')

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    clib.profile_enable(True)
    clib.profile_read()
    launches0 = clib.kernel_launches()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.time()
    ev[0].record(stream)
    for s in range(args.steps):
        step_device()
        ev[s + 1].record(stream)
    barrier()
    t_wall1 = time.time()
    launches = clib.kernel_launches() - launches0
    kernels = clib.profile_read()
    clib.profile_enable(False)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    dev_ms = ev[0].elapsed_time(ev[-1])
    converged = int((status == 0).sum().item())
    sweeps = int(iters.to(torch.int64).sum().item())       # same inputs every step -> same counts

    # ---- end to end: host buffers through the C ABI --------------------------------------
    ext_host = np.ascontiguousarray(exts_np, np.float32)
    R_host = torch.empty((nz, nb, DIM), dtype=torch.float32, pin_memory=True)
    st_host = torch.empty((nz, nb), dtype=torch.int32, pin_memory=True)
    it_host = torch.empty((nz, nb), dtype=torch.int32, pin_memory=True)

    def step_host():
        clib.check_call(lib.ssn_fixed_point_batch(
            sv, nz, nb, n_sites, clib.W_FROM_Z, z_host.data_ptr(), jd, ext_host.ctypes.data, 0, None,
            R_host.data_ptr(), st_host.data_ptr(), it_host.data_ptr(), 0, clib.MEM_HOST, None),
            'ssn_fixed_point_batch(host)')
        return int((st_host == 0).sum().item())

    e2e_steps = max(2, min(args.steps, 5))
    step_host()
    barrier()
    t0 = time.time()
    e2e_conv = sum(step_host() for _ in range(e2e_steps))
    e2e_s = time.time() - t0
    h2d = z_host.numel() * 4 + ext_host.nbytes
    d2h = R_host.numel() * 4 + st_host.numel() * 4 + it_host.numel() * 4

    # ---- end to end through the reference-facing Python API (rank 0) -----------------------------------
    api = {}
    if rank == 0:
        from tc_gan_b200.weight_gen import generate_weight_batch_gpu
        W64 = generate_weight_batch_gpu(n_sites, jds['J'], jds['D'], jds['S'], z_np).astype(np.float64)
        z64 = z_np.astype(np.float64)
        kw = dict(k=P['k'], n=P['n'], r0=np.zeros(DIM))

        def api_call(pairs, **extra):
            t0 = time.time()
            Zs, Rs, info = ssnode.find_fixed_points(nz, iter(pairs), exts_np, method='parallel', **dict(kw, **extra))
            return nz * nb / (time.time() - t0), Rs

        pairs_w = [(z64[i], W64[i]) for i in range(nz)]            # one float64 (z, W) per network, as run/gan.py:622-634
        api_call(pairs_w)
        rate_w, Rs_api = api_call(pairs_w)
        pairs_z = [(z_np[i], None) for i in range(nz)]
        api_call(pairs_z, jds=jds)
        rate_z, _ = api_call(pairs_z, jds=jds)
        api = {'api': {'value': rate_w, 'unit': UNIT, 'call': 'ssnode.find_fixed_points(1024, Z_W_gen, exts): float64 W per network',
                       'h2d_bytes_per_step': nz * DIM * DIM * 4, 'host_bytes_staged_per_step': nz * DIM * DIM * 8 * 2},
               'api_z': {'value': rate_z, 'unit': UNIT, 'call': 'ssnode.find_fixed_points(1024, Z_gen, exts, jds=...): float32 z, W built on the GPU',
                         'h2d_bytes_per_step': nz * DIM * DIM * 4},
               'api_matches_device_path': bool(np.allclose(Rs_api, R.double().cpu().numpy(), rtol=1e-5, atol=1e-5))}
        del W64, z64, pairs_w, pairs_z

    # ---- reduce over ranks (max time, summed work) -------------------------------------------
    stats = torch.tensor([dev_ms, e2e_s, float(converged), float(e2e_conv), float(sweeps)],
                         dtype=torch.float64, device=dev)
    if world > 1:
        tmax = stats.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = stats.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s = tmax[0].item(), tmax[1].item()
        conv_all, e2e_conv_all = tsum[2].item(), tsum[3].item()
    else:
        conv_all, e2e_conv_all = float(converged), float(e2e_conv)

    # ---- secondary block (all ranks take part: the GAN steps shard their networks and all-reduce) ----------
    peak = fp32_peak_tflops((clocks or {}).get('sm_max_mhz')) if rank == 0 else fp32_peak_tflops(None)
    secondary = None
    if not args.no_secondary:
        sec_steps = max(2, min(args.steps, 5))
        secondary = {}
        for wl in ('gan_fp', 'gan_bptt'):
            res = time_gan_step(wl, sec_steps, 3, dev, world, rank, peak)
            secondary[wl] = res
            torch.cuda.empty_cache()
        if rank == 0:
            secondary['nb50'] = time_nb50_slice(dev, peak)

    if rank == 0:
        value = conv_all * args.steps / (dev_ms * 1e-3)
        flops_per_step = sweeps * FLOP_PER_SWEEP          # rank 0's kernel: reference-equivalent sweeps
        k1 = kernels.get('ssn_fp_ws_kernel') or kernels.get('ssn_fp_cluster_kernel')
        kernel_ms = k1[0] / k1[1] if k1 else dev_ms / args.steps
        achieved = flops_per_step / (kernel_ms * 1e-3) * 1e-12
        probe = ctypes.c_double(0.0)
        clib.check_call(lib.ssn_measure_fp32_peak(ctypes.byref(probe)), 'ssn_measure_fp32_peak')
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except (OSError, ValueError):
            pass
        algo_bytes = nz * (DIM * DIM * 4 + nb * DIM * 4 + nb * 8) + nb * DIM * 4
        cs, rc = ctypes.c_int(0), ctypes.c_int(0)
        lib.ssn_fixed_point_occupancy(n_sites, ctypes.byref(cs), ctypes.byref(rc))
        tag = clib.fixed_point_kernel_tag(n_sites)
        per_net, capture = k1_dram_traffic(tag)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dev_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 contraction / f64 state', 'data': 'synthetic',
            'config': workload_config(nz),
            'kernel': {'name': tag, 'cluster_size': cs.value, 'resident_clusters': rc.value,
                       'sms_used': cs.value * rc.value, 'mean_sweeps_per_solve': sweeps / float(nz * nb),
                       'ms_per_launch': kernel_ms, 'share_of_step': kernel_ms * args.steps / dev_ms,
                       'launches': {k: v[1] for k, v in kernels.items()}},
            'clocks': clocks,
            'e2e': dict({'value': e2e_conv_all / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                         'd2h_bytes_per_step': d2h, 'steps': e2e_steps,
                         'call': 'ssn_fixed_point_batch(SSN_MEM_HOST): pinned float32 z in, R/status/iters out'}, **api),
            'gpu_launches': launches,
            'roofline': {'bound': 'fp32_ffma', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': achieved / peak,
                         'traffic': None if per_net is None else per_net * nz,
                         'traffic_source': capture, 'algorithmic_bytes': algo_bytes,
                         'peak_source': 'nominal 148 SM x 128 FMA/clk x 2 x max SM clock (MEASURED_PEAKS.json has no FP32 '
                                        'figure); probe kernel measured %.1f TFLOP/s in this run' % probe.value,
                         'ffma_probe_tflops': probe.value,
                         'hbm': {'achieved_gbs': algo_bytes / (kernel_ms * 1e-3) * 1e-9,
                                 'peak_gbs': peaks.get('hbm_gbs', 6650.0),
                                 'peak_source': 'MEASURED_PEAKS.json' if 'hbm_gbs' in peaks else 'fallback'}},
        }
        if secondary is not None:
            line['secondary'] = secondary
        if world == 1 and not args.no_cpu_baseline:
            base, parity = cpu_baseline_and_parity(R.double().cpu().numpy(), status.cpu().numpy())
            line['cpu_baseline'], line['parity'] = base, parity
            if not parity['ok']:
                print(json.dumps(line))
                raise SystemExit('bench.py: GPU results differ from the reference solver on the shared network list: %r' % parity)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
