#!/usr/bin/env python
"""Headline benchmark of the statevector gate-application path (BASELINE.json metric):
gate-applications/s on a seeded random Clifford+RX circuit, with the achieved fraction of HBM
bandwidth of the fused tile kernel, next to the reference's CPU PyTorch path timed in the same run.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nqubit n] [--depth d]

One *step* = one full pass of the hot path over one synthetic circuit: |0...0> is (re)initialised
on the device and every gate of the circuit is applied.
  N = 1 : the configuration BASELINE.json's metric is quoted on -- 30 qubits, depth 40, complex64 (1 800 gate
          applications, 8 GiB state); `--nqubit 28` is BASELINE config 2.
  N > 1 : the SAME circuit with the high-order qubit index sharded over the N ranks (strong scaling);
          `--nqubit 33 --depth 30` runs the BASELINE config-4 size.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'gate_applications_per_second'
UNIT = 'gate-apps/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200q', choices=['b200q', 'reference'])
    ap.add_argument('--nqubit', type=int, default=0)
    ap.add_argument('--depth', type=int, default=0)
    ap.add_argument('--chunk-bits', type=int, default=0)
    ap.add_argument('--no-fuse', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=15.0, help='time box of the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--config', default='', choices=['', 'c3', 'c5'],
                    help='c3: 30-qubit complex128 QAOA forward + backward; c5: Fock 8 modes x cutoff 10 (one line each)')
    return ap.parse_args()


def workload(args):
    n = args.nqubit or 30
    depth = args.depth or 40
    tag = {(28, 40): 'config2: ', (33, 30): 'config4 size: ', (30, 40): 'metric config: '}.get((n, depth), '')
    name = f'{tag}{n}-qubit random Clifford+RX, depth {depth}, complex64, single B200'
    if args.gpus > 1:
        name += f' -- sharded over {args.gpus} ranks'
    return n, depth, name


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i',
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [nm for i, nm in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in self.rows
                                                          if len(r) > 3 + i)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.rows)}


def cpu_reference_sample(spec, n, seconds, kind_note):
    """The reference's CPU path (oracle/torch_port.py restates it op for op) on a bounded sample of the
    SAME circuit: the first gates that fit the time box, all host threads."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import gates_np
    import torch_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ops = gates_np.lower_spec(spec, n)
    # warm-up on 2 gates (allocator, thread pool)
    torch_port.run_ops(ops, n, max_gates=2)
    _, done, secs = torch_port.run_ops(ops, n, time_budget_s=seconds)
    return {'value': done / secs, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'first {done} of {len(ops)} gates of the same {n}-qubit circuit in {secs:.1f} s '
                      f'({kind_note}; torch CPU, {torch.get_num_threads()} threads)'}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (restated op for op in
    oracle/torch_port.py -- the reference is pure Python and does not exist on the GPU box)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from deepquantum_b200 import workloads as wl
    n, depth, name = workload(args)
    spec = wl.random_clifford_rx_spec(n, depth)
    ngates = wl.count_gates(spec, n)
    per_step = max(2.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_sample(spec, n, per_step, 'reference CPU path')
        if i >= args.warmup:
            vals.append(r)
    v = sum(x['value'] for x in vals) / len(vals)
    base = vals[-1]
    base['value'] = v
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * ngates / v, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'c64', 'data': 'synthetic',
            'config': {'workload': name, 'gates': ngates, 'note': 'each step is a time-boxed sample of the circuit; '
                       'ms_per_step is extrapolated to the full circuit'},
            'cpu_baseline': base,
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def parity_check(args, dev, n=20, depth=40):
    """The same generator at 20 qubits: full final state of the CUDA path (specialised kernels: n >= 20) against the
    CPU oracle in complex128.  (Amplitude parity at 20-26 qubits is in tests/test_gpu_large.py.)"""
    import numpy as np
    import torch

    import deepquantum_b200 as dq
    from deepquantum_b200 import workloads as wl
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import gates_np
    import torch_port
    spec = wl.random_clifford_rx_spec(n, depth)
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, spec)
    cir.to(dev)
    with torch.no_grad():
        out = cir().reshape(-1).cpu().numpy().astype(np.complex128)
    jit = cir._get_program().plan(torch.complex64).jit_status()
    ref, done, _ = torch_port.run_ops(gates_np.lower_spec(spec, n), n, dtype=torch.complex128)
    ref = ref.numpy()
    return {'nqubit': n, 'depth': depth, 'rel_l2_vs_oracle_c128': float(np.linalg.norm(out - ref) / np.linalg.norm(ref)),
            'tolerance': 2e-6, 'specialised_passes': jit['specialised']}


def run_single(args):
    import torch

    import deepquantum_b200 as dq
    from deepquantum_b200 import circuit as circ
    from deepquantum_b200 import engine
    from deepquantum_b200 import workloads as wl

    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    n, depth, name = workload(args)
    circ.PLAN_OPTIONS.update(chunk_bits=args.chunk_bits, fuse=not args.no_fuse)
    spec = wl.random_clifford_rx_spec(n, depth)
    # RX angles are supplied as data (encode=True) so the end-to-end arm moves real inputs host -> device
    cir = dq.QubitCircuit(n)
    angles = []
    for e in spec:
        if e['g'] == 'rx':
            cir.rx(e['w'][0], encode=True)
            angles.append(e['p'][0])
        else:
            wl.apply_spec(cir, [e])
    cir.observable([0], 'z')
    cir.observable([n // 2, n - 1], 'zz')
    cir.to(dev)
    data_host = torch.tensor(angles, dtype=torch.float32).pin_memory()
    prog = cir._get_program()
    ngates = prog.ngates
    plan = prog.plan(torch.complex64)
    n_passes = plan.n_passes
    state_bytes = (2**n) * 8
    bytes_pass = 2 * state_bytes

    with torch.no_grad():
        data_dev = data_host.to(dev)
        state = torch.empty(2**n, dtype=torch.complex64, device=dev)

        # one step = everything cir(data) does on the device: route the angles into the encoder gates, assemble the
        # matrix buffer, re-initialise |0...0>, run every pass (the same region the N > 1 arm times)
        def device_step(ev=None):
            cir.encode(data_dev)
            mats = prog.low.build_matrices(torch.complex64, dev)
            engine.init_basis_(state, n, 1, 0)
            if ev is not None:
                ev[0].record()
            plan.run(state, mats, 1, 0)
            if ev is not None:
                ev[1].record()

        t_jit = time.perf_counter()
        device_step()          # first run: compiles the specialised pass kernels (or loads the cubin cache)
        torch.cuda.synchronize()
        t_jit = time.perf_counter() - t_jit
        jit = plan.jit_status()
        for _ in range(args.warmup):
            device_step()
        torch.cuda.synchronize()
        ev_all = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev_k = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                for _ in range(args.steps)]
        with ClockSampler(0) as clocks:
            ev_all[0].record()
            for k in range(args.steps):
                device_step(ev_k[k])
            ev_all[1].record()
            torch.cuda.synchronize()
        total_ms = ev_all[0].elapsed_time(ev_all[1])
        kern_ms = sum(a.elapsed_time(b) for a, b in ev_k)
        norm = float(engine.norm2(state, n)[0])

        # the same kernel in its HBM-bound regime: ONE gate per pass (what b200q_apply_gate, the evolve_state
        # drop-in, launches) -- reported next to the fused number, which is issue-bound by design
        from deepquantum_b200 import _lib as L
        single = engine.FusedPlan(n, torch.complex64, [L.make_gate(L.GATE_MAT, [n // 2], [], 0, False,
                                                                    L.GATE_REAL | L.GATE_HADAMARD)])
        hmat = (torch.tensor([[1, 1], [1, -1]], dtype=torch.complex64, device=dev) / 2**0.5).reshape(-1)
        for _ in range(3):
            single.run(state, hmat, 1, 0)
        es = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        es[0].record()
        for _ in range(10):
            single.run(state, hmat, 1, 0)
        es[1].record()
        torch.cuda.synchronize()
        single_ms = es[0].elapsed_time(es[1]) / 10
        del state

        # ---- end to end through the public API: pinned host angles -> cir(data) -> expectation -> host
        def e2e_step():
            d = data_host.to(dev, non_blocking=True)
            cir(d)
            return cir.expectation().cpu()

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0

    parity = parity_check(args, dev)
    ms_per_step = total_ms / args.steps
    value = ngates * args.steps / (total_ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    achieved = n_passes * args.steps * bytes_pass / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json'))).get(
            f'jit_pass_c64_{n}q_bytes_per_launch')
    except Exception:
        pass
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'c64',
        'data': 'synthetic',
        'config': {'workload': name, 'gates': ngates, 'passes': n_passes, 'gates_per_pass': ngates / n_passes,
                   'state_bytes': state_bytes,
                   'l2': f'state ({state_bytes / 2**30:g} GiB) is larger than L2 (126 MB): no flush needed',
                   'tile_bytes': 16 << (args.chunk_bits or 12), 'fused': not args.no_fuse, 'norm2_check': norm,
                   'specialised_passes': jit['specialised'], 'generic_passes': n_passes - jit['specialised'],
                   'first_step_s': t_jit,
                   'timed_region': 'encode + matrix assembly + |0..0> init + all passes, per step (CUDA events)',
                   'parity_check': parity},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic,
                     'kernel': ('b200qj_pass (per-pass specialised kernel, NVRTC sm_100a)' if jit['specialised'] else
                                'b200q_tile_kernel<float,12,lean>'),
                     'peak_source': 'MEASURED_PEAKS.json (measured copy, burst)' if peaks else 'fallback 6650',
                     'bytes_per_launch': bytes_pass, 'ms_per_launch': kern_ms / (n_passes * args.steps),
                     'note': 'achieved = passes x bytes_per_launch / time of plan.run (CUDA events around the passes only)',
                     'single_gate_pass': {'ms': single_ms, 'achieved': bytes_pass / (single_ms * 1e-3) / 1e9,
                                          'frac': bytes_pass / (single_ms * 1e-3) / 1e9 / peak}},
        'e2e': {'value': ngates * args.steps / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': data_host.numel() * 4,
                'd2h_bytes_per_step': int(res.numel() * res.element_size()),
                'note': 'cir(data) from pinned host angles + expectation() read back, wall clock'},
        'gpu_launches': args.steps * (n_passes + 1),   # tile kernel per pass + init_basis, timed region only
        'clocks': clocks.summary(),
    }
    if not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_reference_sample(spec, n, args.cpu_seconds, 'port of the reference CPU path')
    print(json.dumps(line))


def run_config(args):
    """BASELINE configs 3 and 5 as bench lines of their own (the default run is the headline metric)."""
    import torch

    import deepquantum_b200 as dq
    from deepquantum_b200 import workloads as wl
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(0) as clocks:
            e0.record()
            for _ in range(args.steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, clocks.summary()

    if args.config == 'c3':
        n, p = args.nqubit or 30, 4
        edges, weights, layout = wl.qaoa_maxcut_structure(n, p)
        cir = dq.QubitCircuit(n)
        wl.build_qaoa(cir, edges, p)
        cir.to('cuda', torch.double)
        params_host = torch.tensor([0.1] * p + [1.0] * p, dtype=torch.float64).pin_memory()
        w = torch.tensor(weights, dtype=torch.float64, device='cuda')

        def step():
            prm = params_host.to('cuda', non_blocking=True).requires_grad_(True)
            cir(wl.qaoa_data(prm, weights, layout))
            loss = 0.5 * (w * (cir.expectation().reshape(-1) - 1)).sum()
            loss.backward()
            return float(loss), prm.grad.cpu()

        ms, clocks = timed(step)
        prog = cir._get_program()
        plan = prog.plan(torch.complex128)
        bytes_pass = 2 * (2**n) * 16
        line = {'metric': 'gate_applications_per_second', 'value': prog.ngates / (ms * 1e-3), 'unit': UNIT, 'n_gpus': 1,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
                'scaling': 'strong', 'vs_baseline': None, 'dtype': 'c128', 'data': 'synthetic',
                'config': {'workload': f'config3: {n}-qubit QAOA MaxCut p={p}, complex128, forward + expectation + backward',
                           'gates': prog.ngates, 'observables': len(edges), 'passes': plan.n_passes,
                           'specialised_passes': plan.jit_status()['specialised'],
                           'max_mem_GiB': torch.cuda.max_memory_allocated() / 2**30},
                'roofline': {'bound': 'hbm', 'achieved': None, 'peak': peak, 'unit': 'GB/s', 'frac': None, 'traffic': None,
                             'note': f'forward: {plan.n_passes} passes of {bytes_pass} bytes; the reverse sweep moves two '
                                     'states per pass (see tools/bench_configs.py for the split)'},
                'e2e': {'value': prog.ngates / (ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': params_host.numel() * 8,
                        'd2h_bytes_per_step': 8 + params_host.numel() * 8,
                        'note': 'the timed step starts from pinned host parameters and reads loss + gradient back'},
                'gpu_launches': args.steps * (2 * plan.n_passes + 4), 'clocks': clocks}
    else:
        nmode, cutoff = 8, 10
        spec = wl.fock_interferometer_spec(nmode)
        cir = dq.QumodeCircuit(nmode, 'vac', cutoff=cutoff, backend='fock', basis=False)
        for e in spec:
            if e['g'] == 's':
                cir.s(e['w'][0], e['p'][0], e['p'][1])
            else:
                cir.bs(e['w'], e['p'])
        cir.to('cuda')
        ms, clocks = timed(lambda: cir())
        bytes_pass = 2 * cutoff**nmode * 8
        passes = cir.fock_plan_stats()['passes']      # a two-mode gate + the one-mode gates around it share a pass
        line = {'metric': 'gate_applications_per_second', 'value': len(spec) / (ms * 1e-3), 'unit': UNIT, 'n_gpus': 1,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
                'scaling': 'strong', 'vs_baseline': None, 'dtype': 'c64', 'data': 'synthetic',
                'config': {'workload': f'config5: Fock {nmode} modes x cutoff {cutoff}, squeezers + Clements mesh, complex64',
                           'gates': len(spec), 'passes': passes,
                           'frac_if_every_gate_were_a_pass': len(spec) * bytes_pass / (ms * 1e-3) / 1e9 / peak},
                'roofline': {'bound': 'hbm', 'achieved': passes * bytes_pass / (ms * 1e-3) / 1e9, 'peak': peak,
                             'unit': 'GB/s', 'frac': passes * bytes_pass / (ms * 1e-3) / 1e9 / peak, 'traffic': None,
                             'kernel': 'qudit_sector_kernel / qudit_sector_staged_kernel / qudit_group_kernel (block-'
                                       'structured register kernels); the step time includes the assembly of the Fock '
                                       'matrices'},
                'e2e': {'value': len(spec) / (ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0,
                        'note': 'parameters live on the device (the reference builds this circuit from constants too)'},
                'gpu_launches': args.steps * 3 * passes, 'clocks': clocks}
    print(json.dumps(line))


def main():
    args = parse()
    if args.config and args.impl != 'reference':
        return run_config(args)
    if args.impl == 'reference':
        return run_reference(args)
    if args.gpus > 1 or int(os.environ.get('WORLD_SIZE', '1')) > 1:
        from deepquantum_b200 import bench_dist
        return bench_dist.run(args)
    return run_single(args)


if __name__ == '__main__':
    main()
