"""GPU micro-measurements used to calibrate the planner cost model (run under gpurun).
Prints one JSON line per experiment."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepquantum_b200 as dq  # noqa: E402
from deepquantum_b200 import _lib as L, engine, workloads as wl  # noqa: E402
from deepquantum_b200 import circuit as circ  # noqa: E402

PEAK = 6547.8e9


def time_plan(plan, st, mats, reps=5, warm=2):
    for _ in range(warm):
        plan.run(st, mats, 1, 0)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        plan.run(st, mats, 1, 0)
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def synthetic(n, tdt, kinds, bits, chunk_bits):
    """kinds: list of 'h' (dense 1q), 'x' (cx), 'd' (diag) applied round-robin over `bits`."""
    gates, off = [], 0
    for i, k in enumerate(kinds):
        b = bits[i % len(bits)]
        if k == 'h':
            gates.append(L.make_gate(L.GATE_MAT, [b], [], 0, False, L.GATE_REAL))
        elif k == 'r':
            gates.append(L.make_gate(L.GATE_MAT, [b], [], 8, False, L.GATE_RXLIKE))
        elif k == 'u':
            gates.append(L.make_gate(L.GATE_MAT, [b], [], 12))
        elif k == 'd':
            gates.append(L.make_gate(L.GATE_DIAG, [b], [], 4))
        else:
            gates.append(L.make_gate(L.GATE_X, [b], [(b + 1) % n], 0))
    return engine.FusedPlan(n, tdt, gates, chunk_bits=chunk_bits)


def main():
    dev = torch.device('cuda')
    out = []
    for tdt, n in ((torch.complex64, 28), (torch.complex128, 27)):
        st = torch.zeros(2**n, dtype=tdt, device=dev)
        st[0] = 1
        h = torch.tensor([[1, 1], [1, -1]], dtype=tdt, device=dev) / 2**0.5
        s = torch.tensor([[1, 0], [0, 1j]], dtype=tdt, device=dev)
        rx = torch.tensor([[0.8, -0.6j], [-0.6j, 0.8]], dtype=tdt, device=dev)
        u = torch.tensor([[0.6, 0.8j], [0.8j, 0.6]], dtype=tdt, device=dev) * (0.6 + 0.8j)
        mats = torch.cat([h.reshape(-1), s.reshape(-1), rx.reshape(-1), u.reshape(-1)])
        bytes_pass = 2 * st.numel() * st.element_size()
        # plain copy for reference
        dst = torch.empty_like(st)
        for _ in range(3):
            dst.copy_(st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dst.copy_(st)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 5
        print(json.dumps({'exp': 'torch_copy', 'dtype': str(tdt), 'n': n, 'ms': t, 'GBps': bytes_pass / t / 1e6}))
        del dst
        hi = [n - 1, n - 2, n - 3, n - 4]          # high bits: single global->global round
        lo = [1, 2, 3, 4]                          # low bits: needs the shared-memory transpose
        mid = [8, 9, 10, 11]
        quick = '--quick' in sys.argv
        for cb in ((12,) if quick else (11, 12, 13)):
            for label, bits in ((('hi', hi), ('lo', lo)) if quick else (('hi', hi), ('lo', lo), ('mid', mid))):
                for kind in ('h', 'r', 'u', 'd', 'x'):
                    for nops in ((1, 8, 32) if quick else (1, 4, 8, 16, 32)):
                        plan = synthetic(n, tdt, [kind] * nops, bits, cb)
                        t = time_plan(plan, st, mats)
                        print(json.dumps({'exp': 'synthetic', 'dtype': str(tdt), 'n': n, 'chunk_bits': cb,
                                          'bits': label, 'kind': kind, 'nops': nops, 'passes': plan.n_passes,
                                          'rounds': plan.stats['n_rounds'], 'ms': round(t, 4),
                                          'GBps_per_pass': round(plan.n_passes * bytes_pass / t / 1e6, 1),
                                          'frac': round(plan.n_passes * bytes_pass / t / 1e-3 / PEAK, 3)}))
                        sys.stdout.flush()
        del st
        torch.cuda.empty_cache()
    # the C2 circuit end to end for each tile size, fused and unfused
    n, depth = 28, 40
    spec = wl.random_clifford_rx_spec(n, depth)
    for cb in ((11, 12) if '--quick' in sys.argv else (11, 12, 13)):
        for fuse in (True, False):
            if not fuse and cb != 12:
                continue
            circ.PLAN_OPTIONS.update(chunk_bits=cb, fuse=fuse)
            cir = dq.QubitCircuit(n)
            wl.apply_spec(cir, spec if fuse else spec[:420])
            cir.to('cuda')
            with torch.no_grad():
                cir()
                torch.cuda.synchronize()
                prog = cir._get_program()
                plan = prog.plan(torch.complex64)
                mats = prog.low.build_matrices(torch.complex64, dev)
                st = cir.state.reshape(-1)
                t = time_plan(plan, st, mats, reps=3, warm=1)
            bytes_pass = 2 * st.numel() * 8
            print(json.dumps({'exp': 'c2', 'chunk_bits': cb, 'fuse': fuse, 'gates': prog.ngates,
                              'passes': plan.n_passes, 'rounds': plan.stats['n_rounds'], 'ms': round(t, 3),
                              'gate_apps_per_s': round(prog.ngates / t * 1e3, 1),
                              'frac': round(plan.n_passes * bytes_pass / t / 1e-3 / PEAK, 3)}))
            sys.stdout.flush()
            del cir, st
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
