"""Histogram of the op codes / rounds of the plan of a workload (host only; parses b200q_plan_export)."""
import collections
import ctypes as C
import sys

import torch

sys.path.insert(0, '.')
import deepquantum_b200 as dq
from deepquantum_b200 import workloads as wl

MAX_OPS, MAX_ROUNDS, MAX_TILE, MAX_Q = 48, 8, 14, 40


class Op(C.Structure):
    _fields_ = [('kind', C.c_uint8), ('slot', C.c_uint8), ('k', C.c_uint8), ('flags', C.c_uint8),
                ('pool_off', C.c_uint16), ('pool_n', C.c_uint16), ('mat_src', C.c_uint32), ('ctrl_reg', C.c_uint32),
                ('ctrl_loc', C.c_uint32), ('dsel_loc', C.c_uint32 * 2), ('ctrl_glob', C.c_uint64),
                ('dsel_glob', C.c_uint64 * 2), ('tk', C.c_uint8 * 4), ('dsel_slot', C.c_uint8 * 2),
                ('code', C.c_uint8), ('tctrl', C.c_uint8), ('arg', C.c_uint8), ('pad', C.c_uint8 * 3),
                ('gate_id', C.c_uint32)]


class Round(C.Structure):
    _fields_ = [('src_global', C.c_uint8), ('dst_global', C.c_uint8), ('direct', C.c_uint8), ('pad', C.c_uint8),
                ('op_begin', C.c_uint16), ('op_end', C.c_uint16), ('slot_bit', C.c_uint8 * 5),
                ('nonreg_bit', C.c_uint8 * 11)]


class Pass(C.Structure):
    _fields_ = [('n_bits', C.c_uint8), ('n_qubits', C.c_uint8), ('tile_bits', C.c_uint8), ('n_rounds', C.c_uint8),
                ('n_ops', C.c_uint8), ('pool_elems', C.c_uint16), ('n_nontile', C.c_uint8), ('layout', C.c_uint8),
                ('lean', C.c_uint8), ('needs_pool', C.c_uint8), ('has_scale', C.c_uint8), ('n_gctrl', C.c_uint8),
                ('gctrl_ops', C.c_uint8 * MAX_OPS), ('tile_phys', C.c_uint8 * MAX_TILE),
                ('nontile_phys', C.c_uint8 * MAX_Q), ('rounds', Round * MAX_ROUNDS), ('ops', Op * MAX_OPS)]


NAMES = {0: 'HAD', 4: 'ROTX', 8: 'ROTY', 12: 'DIAG_R', 16: 'LSWAP', 20: 'XREL', 21: 'X_C1', 22: 'DIAG_T', 23: 'X_LANE',
         24: 'NONE', 25: 'FAST_REAL', 29: 'FAST_RX', 33: 'FAST_GEN', 37: 'MAT1_SLOW', 38: 'X_SLOW', 39: 'DIAG'}


def name(code):
    base = max(k for k in NAMES if k <= code)
    return NAMES[base]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    spec = wl.random_clifford_rx_spec(n, depth)
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, spec)
    prog = cir._get_program()
    plan = prog.plan(torch.complex64)
    raw = plan.export()
    assert len(raw) % C.sizeof(Pass) == 0, (len(raw), C.sizeof(Pass))
    np_ = len(raw) // C.sizeof(Pass)
    hist = collections.Counter()
    rounds = ops = lean = gsel = 0
    for i in range(np_):
        P = Pass.from_buffer_copy(raw[i * C.sizeof(Pass):(i + 1) * C.sizeof(Pass)])
        rounds += P.n_rounds
        ops += P.n_ops
        lean += P.lean
        for o in range(P.n_ops):
            op = P.ops[o]
            nm = name(op.code)
            if nm == 'DIAG_T' and not (op.dsel_loc[0] or op.dsel_loc[1]) and not op.tctrl:
                nm = 'DIAG_T(tile-uniform)'
            if nm == 'XREL' and not op.ctrl_loc:
                nm = 'XREL(tile-uniform)'
            hist[nm] += 1
    print(f'passes {np_} lean {lean} rounds {rounds} ops {ops} gates {prog.ngates}')
    for k, v in hist.most_common():
        print(f'  {k:24s} {v:5d}  {v / np_:5.1f}/pass')


if __name__ == '__main__':
    main()


def lazy_cnot_potential(n=28, depth=40):
    spec = wl.random_clifford_rx_spec(n, depth)
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, spec)
    plan = cir._get_program().plan(torch.complex64)
    raw = plan.export()
    np_ = len(raw) // C.sizeof(Pass)
    tot = free = 0
    for i in range(np_):
        P = Pass.from_buffer_copy(raw[i * C.sizeof(Pass):(i + 1) * C.sizeof(Pass)])
        for r in range(P.n_rounds):
            Rd = P.rounds[r]
            for o in range(Rd.op_begin, Rd.op_end):
                op = P.ops[o]
                if op.code != 21:
                    continue
                tot += 1
                t, c = (op.arg >> 2) + 1, (op.arg & 3) + 1   # amplitude-level slots
                touched = False
                for o2 in range(o + 1, Rd.op_end):
                    q = P.ops[o2]
                    used = q.ctrl_reg
                    if q.kind in (0, 1, 4):
                        used |= 1 << q.slot
                    for j in range(2):
                        if q.dsel_slot[j] != 0xff:
                            used |= 1 << q.dsel_slot[j]
                    if used & ((1 << t) | (1 << c)):
                        touched = True
                        break
                free += not touched
    print(f'X_C1 total {tot}, last-touch-in-round {free}')


if __name__ == '__main__' and len(sys.argv) > 3:
    lazy_cnot_potential()


def dump_rounds(n=28, depth=40, npasses=2):
    spec = wl.random_clifford_rx_spec(n, depth)
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, spec)
    plan = cir._get_program().plan(torch.complex64)
    raw = plan.export()
    for i in range(3, 3 + npasses):
        P = Pass.from_buffer_copy(raw[i * C.sizeof(Pass):(i + 1) * C.sizeof(Pass)])
        print(f'pass {i}: {P.n_rounds} rounds, {P.n_ops} ops')
        for r in range(P.n_rounds):
            Rd = P.rounds[r]
            seq = []
            for o in range(Rd.op_begin, Rd.op_end):
                op = P.ops[o]
                nm = name(op.code)
                s = op.code - max(k for k in NAMES if k <= op.code) if nm in ('HAD', 'ROTX', 'ROTY', 'DIAG_R', 'LSWAP') else (
                    op.arg if nm in ('XREL', 'X_C1') else '')
                seq.append(f'{nm}{s}{"c" if op.tctrl else ""}')
            print(f'  round {r} [{"G" if Rd.src_global else "s"}->{"G" if Rd.dst_global else "s"}] ' + ' '.join(seq))


if __name__ == '__main__' and len(sys.argv) > 3 and sys.argv[3] == 'dump':
    dump_rounds()
