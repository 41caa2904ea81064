#!/usr/bin/env python
"""The reference-side binding of INTEGRATION.md section 2 as an importable file: the ctypes stub a DeepQuantum
maintainer would add as `deepquantum/_b200q.py`.  `install()` replaces `qmath.evolve_state` (qmath.py:485-506),
`Gate.op_state_control` (operation.py:203-219) and `evolve_den_mat` (qmath.py:509-540) of an imported reference
package by calls into libb200q.so for CUDA tensors; CPU tensors keep the reference's own path.

    python tools/patch_reference.py            # import check: installs the stub into the reference (if it is mounted)
                                               # and runs a CPU circuit through the patched entry points

Not imported by the product package.  The library is looked up next to the product package unless B200Q_LIB is set.
"""
import ctypes
import os

import torch

_LIB_PATH = os.environ.get('B200Q_LIB') or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        'deepquantum_b200', 'lib', 'libb200q.so')
_lib = ctypes.CDLL(_LIB_PATH)
_lib.b200q_apply_gate.restype = ctypes.c_int
_lib.b200q_apply_gate.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.POINTER(ctypes.c_int32),
                                  ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
_lib.b200q_qudit_apply.restype = ctypes.c_int
_lib.b200q_qudit_apply.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                   ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_int64, ctypes.c_void_p]
_lib.b200q_last_error.restype = ctypes.c_char_p
_DT = {torch.complex64: 0, torch.complex128: 1}

def _check(rc):
    if rc: raise RuntimeError(_lib.b200q_last_error().decode())

def evolve_state(state, matrix, nqudit, wires, qudit=2):          # replaces qmath.py:485-506
    if not state.is_cuda:
        return _reference_evolve_state(state, matrix, nqudit, wires, qudit)   # the reference keeps its CPU path
    out = state.reshape(-1, qudit ** nqudit).contiguous().clone()  # reference semantics: new tensor, input untouched
    m = matrix.to(out.dtype).contiguous()
    stream = torch.cuda.current_stream(out.device).cuda_stream
    if qudit == 2:
        t = (ctypes.c_int32 * len(wires))(*[nqudit - 1 - w for w in reversed(wires)])   # matrix LSB first
        _check(_lib.b200q_apply_gate(out.data_ptr(), nqudit, _DT[out.dtype], 0, m.data_ptr(), t, len(wires),
                                     None, 0, 0, out.shape[0], 0, stream))
    else:
        w = (ctypes.c_int32 * len(wires))(*wires)
        _check(_lib.b200q_qudit_apply(out.data_ptr(), nqudit, qudit, _DT[out.dtype], m.data_ptr(), w, len(wires),
                                      out.shape[0], stream))
    return out.reshape(state.shape)

def op_state_control(self, x, matrix):                             # replaces operation.py:203-219
    if not x.is_cuda:
        return _reference_op_state_control(self, x, matrix)
    n = self.nqubit
    out = x.reshape(-1, 2 ** n).contiguous().clone()
    t = (ctypes.c_int32 * len(self.wires))(*[n - 1 - w for w in reversed(self.wires)])
    c = (ctypes.c_int32 * len(self.controls))(*[n - 1 - w for w in self.controls])
    _check(_lib.b200q_apply_gate(out.data_ptr(), n, _DT[out.dtype], 0, matrix.to(out.dtype).contiguous().data_ptr(),
                                 t, len(t), c, len(c), 0, out.shape[0], 0,
                                 torch.cuda.current_stream(out.device).cuda_stream))
    return out.reshape(x.shape)

def evolve_den_mat(state, matrix, nqudit, wires, qudit=2):         # replaces qmath.py:509-540
    if not state.is_cuda:
        return _reference_evolve_den_mat(state, matrix, nqudit, wires, qudit)
    # rho [batch, 2, ..., 2] (2n axes) is a 2n-qubit amplitude vector: U on the row wires, conj(U) on the column wires
    out = evolve_state(state, matrix, 2 * nqudit, list(wires), qudit)
    return evolve_state(out, matrix.conj().resolve_conj(), 2 * nqudit, [w + nqudit for w in wires], qudit)

def install(dq=None):
    if dq is None:
        import deepquantum as dq
    global _reference_evolve_state, _reference_op_state_control, _reference_evolve_den_mat
    _reference_evolve_state = dq.qmath.evolve_state
    _reference_op_state_control = dq.operation.Gate.op_state_control
    _reference_evolve_den_mat = dq.qmath.evolve_den_mat
    for mod in (dq.qmath, dq.operation, dq.distributed, dq.gate, dq.photonic.operation, dq.photonic.distributed):
        mod.evolve_state = evolve_state
    for mod in (dq.qmath, dq.operation):
        mod.evolve_den_mat = evolve_den_mat      # Gate.op_den_mat_base and Channel.op_den_mat (the stub loops over the Kraus operators where the reference vmaps)
    dq.operation.Gate.op_state_control = op_state_control


def _self_check():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
    import ref_loader
    if not ref_loader.reference_available():
        print('reference tree not mounted: stub imported, symbols bound, nothing to patch')
        return
    dq = ref_loader.load_reference()
    cir = dq.QubitCircuit(4)
    cir.hlayer(); cir.cnot(0, 1); cir.rx(2, 0.3); cir.toffoli(0, 1, 3)
    before = cir()
    install(dq)
    assert dq.operation.evolve_state is evolve_state and dq.operation.Gate.op_state_control is op_state_control
    after = cir()                      # CPU tensors: the patched entry points hand back to the reference
    assert torch.equal(before, after)
    print('stub installed into the reference; CPU circuit unchanged through the patched entry points')


if __name__ == '__main__':
    _self_check()
