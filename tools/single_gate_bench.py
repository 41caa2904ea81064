"""One un-fused gate per pass on a 28-qubit complex64 state: time and fraction of the HBM roofline."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepquantum_b200 import _lib as L, engine  # noqa: E402

n = 28
dev = torch.device('cuda')
st = torch.zeros(2**n, dtype=torch.complex64, device=dev)
st[0] = 1
h = (torch.tensor([[1, 1], [1, -1]], dtype=torch.complex64, device=dev) / 2**0.5).reshape(-1)
peak = 6551.0
for label, gate in (('H hi', L.make_gate(L.GATE_MAT, [n - 2], [], 0, False, L.GATE_REAL | L.GATE_HADAMARD)),
                    ('H lo', L.make_gate(L.GATE_MAT, [2], [], 0, False, L.GATE_REAL | L.GATE_HADAMARD)),
                    ('H bit0', L.make_gate(L.GATE_MAT, [0], [], 0, False, L.GATE_REAL | L.GATE_HADAMARD)),
                    ('CX hi', L.make_gate(L.GATE_X, [n - 2], [n - 5], 0))):
    plan = engine.FusedPlan(n, torch.complex64, [gate])
    for _ in range(3):
        plan.run(st, h, 1, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        plan.run(st, h, 1, 0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({'gate': label, 'ctas_per_sm': os.environ.get('B200Q_CTAS_PER_SM', 'default'), 'ms': round(ms, 4),
                      'GBps': round(2 * st.numel() * 8 / ms / 1e6, 1), 'frac': round(2 * st.numel() * 8 / ms / 1e6 / peak, 3),
                      'rounds': plan.stats['n_rounds']}))
