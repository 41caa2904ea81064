"""Supplementary timings of BASELINE configs 1, 3 and 5 on one GPU (bench.py is the headline config 2).
One JSON line per config."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepquantum_b200 as dq  # noqa: E402
from deepquantum_b200 import workloads as wl  # noqa: E402

PEAK = 6547.8e9


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def c1():
    cir = dq.QubitCircuit(12)
    wl.apply_spec(cir, wl.c1_plumbing_spec(12))
    cir.to('cuda')
    with torch.no_grad():
        ms = timed(lambda: cir(), reps=20, warm=3)
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(20):
            cir()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 20 * 1e3
    print(json.dumps({'config': 'C1 12q plumbing (35 gates)', 'ms_device': ms, 'ms_wall': wall,
                      'gate_apps_per_s': 35 / wall * 1e3, 'passes': cir._get_program().plan(torch.complex64).n_passes}))


def c3(n=30, p=4):
    edges, weights, layout = wl.qaoa_maxcut_structure(n, p)
    cir = dq.QubitCircuit(n)
    wl.build_qaoa(cir, edges, p)
    cir.to('cuda', torch.double)
    params = torch.tensor([0.1] * p + [1.0] * p, dtype=torch.float64, device='cuda', requires_grad=True)
    w = torch.tensor(weights, dtype=torch.float64, device='cuda')
    prog = cir._get_program()

    def fwd():
        with torch.no_grad():
            cir(wl.qaoa_data(params.detach(), weights, layout))
            return cir.expectation()

    def fwd_bwd():
        params.grad = None
        cir(wl.qaoa_data(params, weights, layout))
        loss = 0.5 * (w * (cir.expectation().reshape(-1) - 1)).sum()
        loss.backward()
        return loss

    ms_f = timed(fwd, reps=2, warm=1)
    ms_fb = timed(fwd_bwd, reps=2, warm=1)
    plan = prog.plan(torch.complex128)
    bytes_pass = 2 * (2**n) * 16
    print(json.dumps({'config': f'C3 QAOA MaxCut p={p}, {n}q complex128', 'gates': prog.ngates, 'observables': len(edges),
                      'passes': plan.n_passes, 'ms_forward_plus_expectation': ms_f, 'ms_forward_backward': ms_fb,
                      'gate_apps_per_s_forward': prog.ngates / ms_f * 1e3,
                      'forward_frac_of_hbm': plan.n_passes * bytes_pass / (ms_f * 1e-3) / PEAK,
                      'loss': float(fwd_bwd()), 'grad': params.grad.tolist(),
                      'max_mem_GiB': torch.cuda.max_memory_allocated() / 2**30}))


def c5(nmode=8, cutoff=10):
    spec = wl.fock_interferometer_spec(nmode)
    cir = dq.QumodeCircuit(nmode, 'vac', cutoff=cutoff, backend='fock', basis=False)
    for e in spec:
        if e['g'] == 's':
            cir.s(e['w'][0], e['p'][0], e['p'][1])
        else:
            cir.bs(e['w'], e['p'])
    cir.to('cuda')
    from deepquantum_b200 import photonic as ph
    ph.FUSE_FOCK = True
    ms = timed(lambda: cir(), reps=5, warm=3)
    st = cir()
    norm = float((st.real**2 + st.imag**2).sum())
    bytes_pass = 2 * cutoff**nmode * 8
    stats = cir.fock_plan_stats()
    ph.FUSE_FOCK = False
    ms_unfused = timed(lambda: cir(), reps=3, warm=1)
    print(json.dumps({'config': f'C5 Fock {nmode} modes cutoff {cutoff} complex64', 'gates': len(spec), 'ms_fused': ms,
                      'passes': stats['passes'], 'gates_per_pass': stats['gates_per_pass'],
                      'ms_one_gate_per_pass': ms_unfused, 'gate_apps_per_s': len(spec) / ms * 1e3,
                      'frac_of_hbm_per_pass_incl_matrix_build': stats['passes'] * bytes_pass / (ms * 1e-3) / PEAK,
                      'norm2': norm}))


if __name__ == '__main__':
    which = sys.argv[1:] or ['c1', 'c3', 'c5']
    if 'c1' in which:
        c1()
    if 'c5' in which:
        c5()
    if 'c3' in which:
        c3()
    if 'c3small' in which:      # profiling size
        c3(n=24)
