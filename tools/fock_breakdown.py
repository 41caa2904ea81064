"""C5 (8 modes, cutoff 10) timing breakdown: Fock-matrix build vs the qudit kernel per gate type."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepquantum_b200 as dq  # noqa: E402
from deepquantum_b200 import workloads as wl  # noqa: E402
from deepquantum_b200.photonic import qudit_apply_  # noqa: E402


def ev():
    return torch.cuda.Event(enable_timing=True)


nmode, cutoff = 8, 10
spec = wl.fock_interferometer_spec(nmode)
cir = dq.QumodeCircuit(nmode, 'vac', cutoff=cutoff, backend='fock', basis=False)
for e in spec:
    if e['g'] == 's':
        cir.s(e['w'][0], e['p'][0], e['p'][1])
    else:
        cir.bs(e['w'], e['p'])
cir.to('cuda')
cir()
torch.cuda.synchronize()
a, b = ev(), ev()
a.record()
for _ in range(5):
    mats = cir.build_matrices(torch.complex64, 'cuda')
b.record()
torch.cuda.synchronize()
print(json.dumps({'build_matrices_ms': a.elapsed_time(b) / 5}))
flat = torch.zeros(cutoff**nmode, dtype=torch.complex64, device='cuda')
flat[0] = 1
bytes_pass = 2 * flat.numel() * 8
for op, m in zip(cir.operators, mats):
    for _ in range(2):
        qudit_apply_(flat, nmode, cutoff, m, op.wires, 1)
    a, b = ev(), ev()
    a.record()
    for _ in range(5):
        qudit_apply_(flat, nmode, cutoff, m, op.wires, 1)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(json.dumps({'gate': type(op).__name__, 'wires': op.wires, 'ms': round(ms, 4),
                      'frac_hbm': round(bytes_pass / ms / 1e6 / 6551.0, 3)}))
