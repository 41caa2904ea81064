"""OpenQASM 3.0 import / export (reference qasm3.py:117-156, 166-472): the exported text
equals the reference's, imported programs run (CPU emulator of the kernel body) to the reference's final state."""
import json
import os

import numpy as np
import pytest
import torch

import deepquantum_b200 as dq
from conftest import GOLDEN
from deepquantum_b200 import workloads as wl
from deepquantum_b200.qasm3 import cir_to_qasm3, qasm3_to_cir
from helpers import emu_run_program


def _g():
    return np.load(os.path.join(GOLDEN, 'qasm3.npz'))


def _state(cir):
    cir.to(torch.double)
    out, _ = emu_run_program(cir._get_program(), cir.nqubit, np.complex128)
    return out[0]


def test_export_equals_the_reference_text():
    g = _g()
    meta = json.loads(str(g['export_spec']))
    cir = dq.QubitCircuit(meta['n'])
    wl.apply_spec(cir, meta['spec'])
    cir.wires_measure = [1, 3]
    assert cir_to_qasm3(cir) == str(g['export_text'])


def test_import_of_the_exported_text():
    g = _g()
    cir = qasm3_to_cir(str(g['export_text']))
    assert cir.wires_measure == [1, 3]
    np.testing.assert_allclose(_state(cir), g['export_reimport_state_c128'], atol=1e-12)


def test_import_definitions_controls_and_powers():
    g = _g()
    cir = qasm3_to_cir(str(g['program']))
    assert cir.nqubit == 5 and cir.wires_measure == list(g['wires_measure'])
    assert len(cir.operators) == int(g['n_ops'])
    # fractional powers go through an eigen-decomposition in complex64 in the reference: 1e-6 there
    np.testing.assert_allclose(_state(cir), g['state_c128'], atol=2e-6)


def test_inverse_modifier_inverts():
    """`inv @` follows the OpenQASM 3 specification here (the reference leaves the gate un-inverted)."""
    src = '''OPENQASM 3.0;
    qubit[2] q;
    gate foo(a) x0, x1 { rx(a) x0; s x1; cx x0, x1; u(a, 0.1, 0.4) x1; }
    h q[0]; h q[1];
    foo(0.3) q[0], q[1];
    inv @ foo(0.3) q[0], q[1];
    t q[0]; inv @ t q[0];
    ctrl @ rz(0.4) q[0], q[1]; ctrl @ inv @ rz(0.4) q[0], q[1];
    pow(-2) @ foo(0.2) q[1], q[0]; pow(2) @ foo(0.2) q[1], q[0];
    inv @ pow(0.5) @ x q[1]; pow(0.5) @ x q[1];
    '''
    out = _state(qasm3_to_cir(src))
    np.testing.assert_allclose(out, np.full(4, 0.5), atol=2e-6)


def test_syntax_variants():
    cir = qasm3_to_cir('OPENQASM 3;\nqubit [ 3 ]  q ;\nbit[3] c;\ncx q[0],q[1];\nrx(pi/2)q[2];\nctrl@x q[0],q[2];\n'
                       ' rz ( 0.1 + 0.2 )  q[1] ;\n pow ( 2 ) @ s q[2];\nc = measure q;\n')
    got = [(type(o).__name__, o.wires, o.controls) for o in cir.operators]
    assert got == [('CNOT', [0, 1], []), ('Rx', [2], []), ('PauliX', [2], [0]), ('Rz', [1], []), ('SGate', [2], []),
                   ('SGate', [2], [])]
    assert abs(float(cir.operators[1].theta) - np.pi / 2) < 1e-6 and abs(float(cir.operators[3].theta) - 0.3) < 1e-6
    assert cir.wires_measure == [0, 1, 2]
    bell = qasm3_to_cir('OPENQASM 3.0; qubit[3] q; gate bell a, b { h a; cx a, b; } bell q[0], q[1]; '
                        'ctrl @ bell q[2], q[0], q[1];')
    assert [(type(o).__name__, o.wires, o.controls) for o in bell.operators] == [
        ('Hadamard', [0], []), ('CNOT', [0, 1], []), ('Hadamard', [0], [2]), ('PauliX', [1], [2, 0])]


def test_barrier_comments_and_errors():
    src = '''OPENQASM 3.0;   // header
    include "stdgates.inc";
    /* block
       comment */
    qubit[3] q;
    h q[0]; barrier q[0], q[2]; cx q[0], q[2];
    '''
    cir = qasm3_to_cir(src)
    assert [type(o).__name__ for o in cir.operators] == ['Hadamard', 'Barrier', 'CNOT']
    assert cir.operators[1].wires == [0, 2]
    with pytest.raises(ValueError):
        qasm3_to_cir('qubit[2] q; h q[0];')
    with pytest.raises(ValueError):
        qasm3_to_cir('OPENQASM 3.0; h q[0];')
    with pytest.raises(ValueError):
        qasm3_to_cir('OPENQASM 3.0; qubit[1] q; frobnicate q[0];')
    with pytest.raises(ValueError):
        qasm3_to_cir('OPENQASM 3.0; qubit[1] q; rx(__import__("os")) q[0];')


@pytest.mark.gpu
@pytest.mark.parametrize('rdtype', [torch.float64, torch.float32])
def test_gpu_imported_programs_match_the_reference(rdtype):
    """The same two fixtures through the product path on the device: the imported circuits run by the fused / specialised
    kernels give the reference's final states (complex128 and complex64)."""
    g = _g()
    tol = 1e-10 if rdtype == torch.float64 else 3e-6
    cir = qasm3_to_cir(str(g['export_text']))
    cir.to('cuda', rdtype)
    out = cir().reshape(-1).cpu().numpy()
    assert np.abs(out - g['export_reimport_state_c128']).max() < tol
    cir = qasm3_to_cir(str(g['program']))
    cir.to('cuda', rdtype)
    out = cir().reshape(-1).cpu().numpy()
    assert np.abs(out - g['state_c128']).max() < max(tol, 2e-6)     # fractional powers: 1e-6 in the reference itself
