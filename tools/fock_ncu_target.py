"""Two launches for ncu: a beamsplitter on modes (3, 4) (direct register kernel) and on (6, 7) (staged variant) at
config-5 size.  `ncu --set full -k regex:qudit_sector -c 2 python tools/fock_ncu_target.py`"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepquantum_b200 import photonic as ph  # noqa: E402

n, d = 8, 10
st = torch.randn(1, d**n, dtype=torch.complex64, device='cuda')
for w in ((3, 4), (6, 7)):
    op = ph.BeamSplitter([0.3, 1.0], n, list(w), d)
    m = op.update_matrix_state().reshape(d * d, -1).to(torch.complex64).to('cuda')
    ph.qudit_apply_(st, n, d, m, op.wires, 1, op._structure)
torch.cuda.synchronize()
