"""Parity of the CUDA path (through the C ABI / the product API) with the CPU oracle and the
reference-generated fixtures.  Run on the B200 box: pytest -m gpu."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

import gates_np
import statevec_oracle as so
from conftest import GOLDEN
from helpers import lower_ops

import deepquantum_b200 as dq
from deepquantum_b200 import _lib as L
from deepquantum_b200 import engine
from deepquantum_b200 import workloads as wl

pytestmark = pytest.mark.gpu

# tolerances: complex128 engine vs complex128 reference rel-L2 <= 1e-10;
# complex64 engine vs complex128 reference rel-L2 <= 2e-6 at depth <= 40 (the reference's own complex64
# path sits at 6-8e-7 there).
TOL = {np.complex128: 1e-10, np.complex64: 2e-6}


def _golden(name):
    return np.load(os.path.join(GOLDEN, name))


def _cases():
    g = _golden('circuits.npz')
    return sorted({k.split('/')[0] for k in g.files if k.endswith('/spec') and not k.startswith('batched')})


def _run_ops_gpu(ops, n, cdtype, state=None, batch=1, **opts):
    arr, ng, mats = lower_ops(ops, n, cdtype)
    tdt = torch.complex64 if cdtype == np.complex64 else torch.complex128
    if state is None:
        st = torch.zeros(batch, 2**n, dtype=tdt, device='cuda')
        st[:, 0] = 1
    else:
        st = torch.tensor(np.asarray(state).reshape(batch, 2**n), dtype=tdt, device='cuda').contiguous()
    plan = engine.FusedPlan(n, tdt, list(arr)[:ng], **opts)
    plan.run(st, torch.tensor(mats, device='cuda'), batch, 0)
    torch.cuda.synchronize()
    return st.cpu().numpy(), plan


@pytest.mark.parametrize('case', _cases())
@pytest.mark.parametrize('rdtype', [torch.float64, torch.float32])
def test_golden_circuits_product_api(case, rdtype):
    g = _golden('circuits.npz')
    meta = json.loads(str(g[case + '/spec']))
    n, spec = meta['n'], meta['spec']
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, spec, torch.complex128 if rdtype == torch.float64 else torch.complex64)
    cir.to('cuda', rdtype)
    out = cir().reshape(-1).cpu().numpy()
    ref = g[case + '/c128']
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert err < (1e-10 if rdtype == torch.float64 else 2e-6), err
    if rdtype == torch.float32:   # and against the reference's own complex64 result
        assert np.linalg.norm(out - g[case + '/c64']) / np.linalg.norm(ref) < 3e-6


@pytest.mark.parametrize('chunk_bits', [11, 12, 13])
@pytest.mark.parametrize('fuse', [True, False])
@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
def test_every_bit_position(chunk_bits, fuse, cdtype):
    n = 16
    rng = np.random.default_rng(5)
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi /= np.linalg.norm(psi)
    ops = []
    for w in range(n):
        ops.append((gates_np.u3(*rng.uniform(0, 6, 3)), [w], []))
        ops.append((gates_np.X, [w], [(w + 3) % n, (w + 7) % n]))
        ops.append((gates_np.rz(0.7 + w), [w], []))
        ops.append((gates_np.rzz(0.2 + w), [w, (w + 5) % n], [(w + 1) % n]))
        ops.append((gates_np.ry(0.4 + w), [(w + 2) % n], [w]))
    for k, wires, ctr in [(2, [n - 1, 2], []), (2, [0, n - 2], [5]), (3, [1, n - 1, 6], []), (4, [3, 0, n - 3, 8], [1])]:
        q, _ = np.linalg.qr(rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k)))
        ops.append((q, wires, ctr))
    ref = so.run_circuit(ops, n, state=psi)
    out, plan = _run_ops_gpu(ops, n, cdtype, state=psi, chunk_bits=chunk_bits, fuse=fuse)
    err = np.linalg.norm(out[0] - ref)
    assert err < TOL[cdtype], (err, plan.stats)
    assert plan.n_passes == (len(ops) if not fuse else plan.n_passes)


@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
@pytest.mark.parametrize('n', [1, 2, 3, 5, 8])
def test_tiny_states(n, cdtype):
    rng = np.random.default_rng(n)
    ops = []
    for _ in range(10):
        w = int(rng.integers(n))
        ops.append((gates_np.u3(*rng.uniform(0, 6, 3)), [w], []))
        if n > 1:
            c = int((w + 1 + rng.integers(n - 1)) % n)
            ops.append((gates_np.X, [w], [c]))
            ops.append((gates_np.rz(0.3), [c], [w]))
    ref = so.run_circuit(ops, n)
    out, _ = _run_ops_gpu(ops, n, cdtype)
    assert np.linalg.norm(out[0] - ref) < TOL[cdtype]


def test_batched_initial_states():
    g = _golden('circuits.npz')
    meta = json.loads(str(g['batched_n6/spec']))
    n = meta['n']
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, meta['spec'])
    cir.to('cuda', torch.double)
    init = torch.tensor(g['batched_n6/init'], device='cuda').unsqueeze(-1)
    out = cir(state=init)
    assert out.shape == (3, 2**n, 1)
    np.testing.assert_allclose(out.squeeze(-1).cpu().numpy(), g['batched_n6/c128'], atol=1e-12)


def test_batched_data():
    n = 5
    cir = dq.QubitCircuit(n)
    cir.hlayer()
    cir.rxlayer(encode=True)
    cir.cnot_ring()
    cir.rzz([0, 3], encode=True)
    cir.to('cuda', torch.double)
    data = torch.rand(4, 6, dtype=torch.float64, device='cuda')
    out = cir(data)
    assert out.shape == (4, 2**n, 1)
    for b in range(4):
        ops = [(gates_np.H, [q], []) for q in range(n)]
        ops += [(gates_np.rx(float(data[b, q])), [q], []) for q in range(n)]
        ops += [(gates_np.CNOT, [q, (q + 1) % n], []) for q in range(n)]
        ops += [(gates_np.rzz(float(data[b, 5])), [0, 3], [])]
        ref = so.run_circuit(ops, n)
        np.testing.assert_allclose(out[b, :, 0].cpu().numpy(), ref, atol=1e-12)


def test_evolve_state_boundary_noncontiguous():
    """qmath.evolve_state semantics: any strides in, same shape out, input untouched."""
    n = 7
    rng = np.random.default_rng(3)
    psi = rng.normal(size=(2, 2**n)) + 1j * rng.normal(size=(2, 2**n))
    u = gates_np.u3(0.3, 0.9, -1.2)
    x = torch.tensor(psi, device='cuda').reshape([2] + [2] * n).permute(0, 3, 1, 2, 4, 5, 6, 7)
    keep = x.clone()
    # undo the permutation logically: wire order of x is (2,0,1,3,...)
    y = dq.evolve_state(x, torch.tensor(u, device='cuda'), n, [1])
    assert y.shape == x.shape and torch.equal(x, keep)
    ref = so.evolve_state(np.ascontiguousarray(keep.cpu().numpy()).reshape(2, -1), u, n, [1])
    np.testing.assert_allclose(y.reshape(2, -1).cpu().numpy(), ref, atol=1e-13)
    q, _ = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))
    y2 = dq.evolve_state(torch.tensor(psi, device='cuda').reshape([2] + [2] * n), torch.tensor(q, device='cuda'), n,
                         [5, 2])
    np.testing.assert_allclose(y2.reshape(2, -1).cpu().numpy(), so.evolve_state(psi, q, n, [5, 2]), atol=1e-13)


def test_gate_modules_standalone():
    """`dq.Rx(theta)(state)` style use (tutorials/basics.ipynb cells 9-18 of the reference)."""
    one = torch.tensor([0, 1], dtype=torch.cfloat, device='cuda')
    out = dq.PauliX().to('cuda')(one)
    np.testing.assert_allclose(out.reshape(-1).cpu().numpy(), [1, 0], atol=1e-7)
    out = dq.Rx(torch.pi / 2).to('cuda')(one)
    np.testing.assert_allclose(out.reshape(-1).cpu().numpy(), [-0.70710678j, 0.70710678], atol=1e-6)
    st = torch.zeros(4, dtype=torch.cfloat, device='cuda')
    st[0] = 1
    out = dq.Swap(nqubit=2, wires=[0, 1]).to('cuda')(dq.PauliX(nqubit=2, wires=[1]).to('cuda')(st))
    np.testing.assert_allclose(out.reshape(-1).cpu().numpy(), [0, 0, 1, 0], atol=1e-7)


def test_expectation_pauli_strings():
    n = 9
    spec = wl.random_clifford_rx_spec(n, 6, seed=21)
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, spec)
    obs = [([0], 'z'), ([1, 5], 'zz'), ([2, 3, 8], 'xyz'), ([4], 'y'), ([7, 6], 'xx'), ([0, 4, 8], 'zzz')]
    for w, b in obs:
        cir.observable(w, b)
    cir.to('cuda', torch.double)
    psi = cir().reshape(-1).cpu().numpy()
    ref = so.run_circuit(gates_np.lower_spec(spec, n), n)
    assert np.linalg.norm(psi - ref) < 1e-10
    exp = cir.expectation().cpu().numpy()
    want = [so.expectation_pauli(ref, n, w, b) for w, b in obs]
    np.testing.assert_allclose(exp, want, atol=1e-10)


def test_reductions():
    n = 18
    rng = np.random.default_rng(1)
    a = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    b = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    for tdt, tol in ((torch.complex128, 1e-12), (torch.complex64, 1e-5)):
        ta, tb = torch.tensor(a, dtype=tdt, device='cuda'), torch.tensor(b, dtype=tdt, device='cuda')
        assert abs(float(engine.norm2(ta, n)[0]) / np.vdot(a, a).real - 1) < tol
        ip = complex(engine.inner_product(ta, tb, n)[0])
        assert abs(ip - np.vdot(a, b)) / abs(np.vdot(a, b)) < max(tol, 1e-4 if tdt == torch.complex64 else 0)


@pytest.mark.parametrize('rdtype,n', [(torch.float32, 26), (torch.float64, 25)])
def test_large_state_properties(rdtype, n):
    """Full-size behaviour through size-independent properties: norm preservation and the
    circuit + inverse round trip back to |0...0> (exact inverses only: rotations and CX)."""
    g = torch.Generator().manual_seed(3)
    cir = dq.QubitCircuit(n)
    for _ in range(3):
        ang = (torch.rand(n, 3, generator=g) * 6).tolist()
        for q in range(n):
            cir.u3(q, ang[q])
        perm = torch.randperm(n, generator=g).tolist()
        for i in range(0, n - 1, 2):
            cir.cx(perm[i], perm[i + 1])
        cir.rzz([perm[0], perm[-1]], 0.3)
    cir.to('cuda', rdtype)
    psi = cir()
    tol = 1e-4 if rdtype == torch.float32 else 1e-11
    assert abs(float(engine.norm2(psi.reshape(-1), n)[0]) - 1) < tol
    full = cir + cir.inverse()
    back = full().reshape(-1)
    assert abs(abs(complex(back[0])) - 1) < tol
    assert float(engine.norm2(back, n)[0]) - abs(complex(back[0]))**2 < tol


def test_qaoa_loss_and_gradient_match_reference_autograd():
    """Config 3 shape at small n: forward + expectation + loss.backward() through the adjoint sweep kernel
    against the reference's own autograd result (tests/golden/qaoa.npz)."""
    g = _golden('qaoa.npz')
    for key in sorted({k.split('/')[0] for k in g.files}):
        meta = json.loads(str(g[key + '/meta']))
        n, p, edges, weights = meta['n'], meta['p'], meta['edges'], meta['weights']
        _, _, layout = wl.qaoa_maxcut_structure(n, p, seed=meta['seed'])
        cir = dq.QubitCircuit(n)
        wl.build_qaoa(cir, [tuple(e) for e in edges], p)
        cir.to('cuda', torch.double)
        params = torch.tensor(g[key + '/params'], device='cuda', requires_grad=True)
        data = wl.qaoa_data(params, weights, layout)
        state = cir(data)
        np.testing.assert_allclose(state.detach().reshape(-1).cpu().numpy(), g[key + '/state'], atol=1e-10)
        exp = cir.expectation()
        np.testing.assert_allclose(exp.detach().cpu().numpy(), g[key + '/expectation'], atol=1e-10)
        w = torch.tensor(weights, dtype=torch.float64, device='cuda')
        loss = 0.5 * (w * (exp.reshape(-1) - 1)).sum()
        loss.backward()
        np.testing.assert_allclose(float(loss), float(g[key + '/loss']), atol=1e-10)
        np.testing.assert_allclose(params.grad.cpu().numpy(), g[key + '/grad'], atol=1e-7, rtol=1e-7)


@pytest.mark.parametrize('with_h', [False, True])
@pytest.mark.parametrize('rdtype', [torch.float64, torch.float32])
def test_state_autograd_matches_dense_torch(rdtype, with_h):
    """Autograd of an arbitrary real function of the output state w.r.t. gate parameters AND the input state.

    The backward pass un-computes the state with U^dagger, which inverts U only as far as U is unitary.  With
    exactly unitary gates the complex128 gradient matches dense autograd to 1e-9; the reference's Hadamard is a
    float32-rounded constant even in its complex128 path (gate.py:1069, unitary to 6e-8 only), which bounds
    the gradient accuracy at ~1e-6 -- the same floor as the reference's own adjoint (adjoint.py:60)."""
    import torch_port
    n = 12
    g = torch.Generator().manual_seed(5)

    def first_layer(c):
        if with_h:
            c.hlayer()
        else:
            c.rzlayer()

    cir = dq.QubitCircuit(n)
    first_layer(cir)
    cir.rxlayer()
    cir.cnot_ring()
    cir.u3layer()
    cir.rzz([0, 5])
    cir.rxx([n - 1, 3])
    cir.crz(2, 7)
    cir.cx(4, 9)
    cir.rylayer()
    cir.to('cuda', rdtype)
    cdt = torch.complex128 if rdtype == torch.float64 else torch.complex64
    psi0 = torch.randn(2**n, generator=g, dtype=torch.float64) + 1j * torch.randn(2**n, generator=g,
                                                                                   dtype=torch.float64)
    psi0 = (psi0 / psi0.norm()).to(cdt).cuda().requires_grad_(True)
    wvec = (torch.randn(2**n, generator=g, dtype=torch.float64) + 1j * torch.randn(2**n, generator=g,
                                                                                   dtype=torch.float64)).to(cdt).cuda()
    out = cir(state=psi0.reshape(-1, 1)).reshape(-1)
    loss = (wvec.conj() * out).sum().real + (out.real**2 * torch.arange(2**n, device='cuda') / 2**n).sum()
    loss.backward()
    got = {name: p.grad.detach().cpu().double() for name, p in cir.named_parameters()}
    got_psi = psi0.grad.detach().cpu()
    # dense reference on CPU: same matrices through the port of the reference contraction
    prog = cir._get_program()
    params = {name: p.detach().cpu().double().requires_grad_(True) for name, p in cir.named_parameters()}
    cpu = dq.QubitCircuit(n)
    first_layer(cpu)
    cpu.rxlayer(); cpu.cnot_ring(); cpu.u3layer(); cpu.rzz([0, 5]); cpu.rxx([n - 1, 3])
    cpu.crz(2, 7); cpu.cx(4, 9); cpu.rylayer()
    cpu.to(torch.double)
    with torch.no_grad():
        for (name, p), (_, q) in zip(cpu.named_parameters(), cir.named_parameters()):
            p.copy_(q.detach().cpu().double())
    x0 = psi0.detach().cpu().to(torch.complex128).requires_grad_(True)
    x = x0.reshape([1] + [2] * n)
    for op in cpu.operators:
        m = op.update_matrix().to(torch.complex128)
        if isinstance(op, dq.CNOT):
            x = torch_port.evolve_state(x, m, n, op.wires)
        elif op.controls:
            x = torch_port.evolve_state_controlled(x, m, n, op.wires, op.controls)
        else:
            x = torch_port.evolve_state(x, m, n, op.wires)
    o = x.reshape(-1)
    ref_loss = (wvec.cpu().to(torch.complex128).conj() * o).sum().real + (o.real**2 * torch.arange(2**n) / 2**n).sum()
    ref_loss.backward()
    tol = (2e-6 if with_h else 1e-9) if rdtype == torch.float64 else 3e-4
    assert abs(float(loss) - float(ref_loss)) < max(tol * 10, 1e-8)
    for (name, p) in cpu.named_parameters():
        assert abs(float(p.grad) - float(got[name])) < tol * max(1.0, abs(float(p.grad))), name
    assert (got_psi.to(torch.complex128) - x0.grad).norm() / x0.grad.norm() < tol


@pytest.mark.parametrize('key', ['m2_c5', 'm3_c4', 'm4_c6', 'm5_c8'])
@pytest.mark.parametrize('rdtype', [torch.float64, torch.float32])
def test_fock_tensor_path(key, rdtype):
    """Config 5 shape: squeezers + beamsplitter mesh + phase shifter on the Fock tensor, against the reference."""
    g = _golden('fock.npz')
    meta = json.loads(str(g[key + '/spec']))
    n, d = meta['nmode'], meta['cutoff']
    cir = dq.QumodeCircuit(n, 'vac', cutoff=d, backend='fock', basis=False)
    for e in meta['spec']:
        if e['g'] == 's':
            cir.s(e['w'][0], e['p'][0], e['p'][1])
        elif e['g'] == 'bs':
            cir.bs(e['w'], e['p'])
        else:
            cir.ps(e['w'][0], e['p'][0])
    cir.to('cuda', rdtype)
    out = cir().reshape(-1).cpu().numpy()
    ref = g[key + '/c128']
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert err < (1e-10 if rdtype == torch.float64 else 2e-6), err
    # the qudit boundary function: evolve_state(..., qudit=cutoff)
    rng = np.random.default_rng(0)
    psi = rng.normal(size=(2, d**n)) + 1j * rng.normal(size=(2, d**n))
    m = rng.normal(size=(d * d, d * d)) + 1j * rng.normal(size=(d * d, d * d))
    y = dq.evolve_state(torch.tensor(psi, device='cuda').reshape([2] + [d] * n), torch.tensor(m, device='cuda'), n,
                        [n - 1, 0], d)
    np.testing.assert_allclose(y.reshape(2, -1).cpu().numpy(), so.evolve_state(psi, m, n, [n - 1, 0], d), atol=1e-10)


def test_sharded_two_gpus():
    """The sharded path over NCCL on 2 GPUs against the single-GPU engine (skipped on a 1-GPU box)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29571', os.path.join(root, 'tests', 'dist_gpu_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'SHARDED_OK' in out.stdout


def test_sharded_fock_two_gpus():
    """The sharded Fock tensor path (reference photonic/distributed.py:65-78) over NCCL on 2 GPUs
    against the host oracle and the single-GPU circuit (skipped on a 1-GPU box)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29573', os.path.join(root, 'tests', 'dist_fock_gpu_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'FOCK_SHARDED_OK' in out.stdout, out.stdout[-2000:]


@pytest.mark.parametrize('structured', [True, False])
@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
def test_structured_and_general_op_codes(cdtype, structured):
    """The lean op set (Hadamard add/sub, three-shear rotations, shear phases, lane transpositions, xor swaps)
    and the general fallbacks of the non-lean kernel, on the same circuit, against the oracle: rotations over
    the full 4*pi period, gates on index bit 0 (the complex64 lane), thread-level and register-slot controls,
    1- and 2-selector diagonals."""
    n = 17
    rng = np.random.default_rng(23)
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi /= np.linalg.norm(psi)
    ops = []
    for _ in range(220):
        w = int(rng.integers(n))
        c = int((w + 1 + rng.integers(n - 1)) % n)
        if rng.integers(5) == 0:
            w, c = n - 1, int(rng.integers(n - 1))
        elif rng.integers(5) == 0:
            c, w = n - 1, int(rng.integers(n - 1))
        th = float(rng.uniform(0, 4 * np.pi))
        kind = int(rng.integers(10))
        if kind == 0:
            ops.append((gates_np.H, [w], []))
        elif kind == 1:
            ops.append((gates_np.rx(th), [w], []))
        elif kind == 2:
            ops.append((gates_np.ry(th), [w], []))
        elif kind == 3:
            ops.append((gates_np.X, [w], [c]))
        elif kind == 4:
            ops.append((gates_np.rx(th), [w], [c]))
        elif kind == 5:
            ops.append((gates_np.S, [w], []))
        elif kind == 6:
            ops.append((gates_np.rz(th), [w], [c] if rng.integers(2) else []))
        elif kind == 7:
            ops.append((gates_np.rzz(th), [c, w], []))
        elif kind == 8:
            ops.append((gates_np.H, [w], [c]))
        else:
            ops.append((gates_np.rx(th).conj().T, [w], []))
    ref = so.run_circuit(ops, n, psi)
    out, plan = _run_ops_gpu(ops, n, cdtype, state=psi, structured=structured)
    err = np.linalg.norm(out[0] - ref) / np.linalg.norm(ref)
    assert err < TOL[cdtype], (err, plan.stats)


def test_batched_data_gradient_matches_per_sample_and_dense():
    """Mini-batch training (2-D data, the reference's vmap path circuit.py:227-241): loss.backward() through the
    reverse sweep for a batch = sum of the single-sample gradients, and = dense torch autograd on the host."""
    import torch_port
    n, nb = 6, 3
    g = torch.Generator().manual_seed(9)

    def build():
        c = dq.QubitCircuit(n)
        c.rxlayer(encode=True)
        c.cnot_ring()
        c.rylayer()
        c.rzz([0, 3])
        c.rxlayer()
        for q in range(n):
            c.observable([q], 'z')
        return c

    cir = build()
    cir.to('cuda', torch.double)
    data = (torch.rand(nb, n, generator=g, dtype=torch.float64) * 3).cuda()
    wts = torch.rand(nb, n, generator=g, dtype=torch.float64).cuda()
    cir(data)
    loss = (cir.expectation().reshape(nb, n) * wts).sum()
    loss.backward()
    got = [p.grad.detach().clone() for p in cir.parameters()]
    # per sample
    acc = [torch.zeros_like(x) for x in got]
    tot = 0.0
    for b in range(nb):
        cir.zero_grad()
        cir(data[b])
        l_b = (cir.expectation().reshape(n) * wts[b]).sum()
        l_b.backward()
        tot += float(l_b)
        for a, p in zip(acc, cir.parameters()):
            a += p.grad
    assert abs(tot - float(loss)) < 1e-10
    for a, b_ in zip(acc, got):
        assert (a - b_).abs().max() < 1e-9, (a, b_)
    # dense autograd on the host for sample 0 .. nb-1 (gate by gate like the reference)
    params = [p.detach().cpu().clone().requires_grad_(True) for p in cir.parameters()]

    def rot(kind, t):
        c, s = torch.cos(t / 2).to(torch.complex128), torch.sin(t / 2).to(torch.complex128)
        if kind == 'rx':
            return torch.stack([c, -1j * s, -1j * s, c]).reshape(2, 2)
        return torch.stack([c, -s, s, c]).reshape(2, 2)

    xm = torch.tensor(gates_np.X, dtype=torch.complex128)
    assert len(params) == 2 * n + 1   # one 0-d theta per gate: n Ry, one Rzz, n Rx
    ry_p, rzz_p, rx_p = params[:n], params[n:n + 1], params[n + 1:]
    loss_ref = 0
    idx = torch.arange(2**n)
    for b in range(nb):
        st = torch.zeros(2**n, dtype=torch.complex128)
        st[0] = 1
        x = st.reshape([1] + [2] * n)
        d = data[b].cpu()
        for q in range(n):
            x = torch_port.evolve_state(x, rot('rx', d[q]), n, [q])
        for q in range(n):
            x = torch_port.evolve_state_controlled(x, xm, n, [(q + 1) % n], [q])
        for q in range(n):
            x = torch_port.evolve_state(x, rot('ry', ry_p[q]), n, [q])
        e = torch.exp(-0.5j * rzz_p[0].to(torch.complex128))
        zz = torch.diag(torch.stack([e, e.conj(), e.conj(), e]))
        x = torch_port.evolve_state(x, zz, n, [0, 3])
        for q in range(n):
            x = torch_port.evolve_state(x, rot('rx', rx_p[q]), n, [q])
        prob = (x.reshape(-1).conj() * x.reshape(-1)).real
        for q in range(n):
            sign = 1 - 2 * ((idx >> (n - 1 - q)) & 1).to(torch.float64)
            loss_ref = loss_ref + wts[b, q].cpu() * (prob * sign).sum()
    loss_ref.backward()
    assert abs(float(loss_ref) - float(loss)) < 1e-9
    for p_ref, g_gpu in zip(params, got):
        assert (p_ref.grad - g_gpu.cpu().reshape(p_ref.shape)).abs().max() < 1e-8


def test_get_unitary_matches_reference():
    """`QubitCircuit.get_unitary()` (reference circuit.py:467-477) against the reference's own result
    (tests/golden/unitary.npz): the fused plan applied to the identity as a batch of 2^n basis states."""
    g = _golden('unitary.npz')
    for case in sorted({k.split('/')[0] for k in g.files}):
        meta = json.loads(str(g[case + '/spec']))
        n, spec = meta['n'], meta['spec']
        cir = dq.QubitCircuit(n)
        wl.apply_spec(cir, spec, torch.complex128)
        cir.to('cuda', torch.double)
        u = cir.get_unitary().cpu().numpy()
        assert np.abs(u - g[case + '/unitary']).max() < 1e-10, case


def test_package_level_measure_and_expectation():
    """`dq.measure` / `dq.expectation` (reference __init__.py:112, qmath.py:568-638, 830-860)."""
    n = 5
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, wl.random_clifford_rx_spec(n, 3, seed=4))
    cir.observable([0, 2], 'zx')
    cir.to('cuda', torch.double)
    st = cir()
    e = dq.expectation(st, cir.observables[0])
    assert e.shape == () and abs(float(e) - float(cir.expectation()[0])) < 1e-14
    psi = st.reshape(-1).cpu().numpy()
    want = so.expectation_pauli(psi, n, [0, 2], 'zx')
    assert abs(float(e) - want) < 1e-10
    res = dq.measure(st, shots=200, with_prob=True)
    assert sum(c for c, _ in res.values()) == 200
    for key, (_, p) in res.items():
        assert abs(p - abs(psi[int(key, 2)])**2) < 1e-12


@pytest.mark.parametrize('n', [12, 21])
@pytest.mark.parametrize('rdtype', [torch.float64, torch.float32])
def test_uany_on_five_and_six_wires(n, rdtype):
    """`cir.any(U, wires=[...])` with 5 and 6 wires (reference gate.py:2745-2790 accepts any number): dense passes of
    their own (b200q_dense_kernel) between fused passes, also controlled and inverted, against the oracle."""
    rng = np.random.default_rng(n)

    def rand_u(k):
        q, _ = np.linalg.qr(rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k)))
        return q

    cdt = torch.complex128 if rdtype == torch.float64 else torch.complex64
    u5, u6 = rand_u(5), rand_u(6)
    w5, w6 = [0, n - 1, 3, 7, 5], [n - 3, 1, 4, 8, 2, 6]
    cir = dq.QubitCircuit(n)
    cir.hlayer()
    cir.any(torch.tensor(u5, dtype=cdt), wires=w5)
    cir.rxlayer(inputs=[0.3 + 0.1 * w for w in range(n)])
    cir.cnot(9, 2)
    cir.any(torch.tensor(u6, dtype=cdt), wires=w6, controls=[10])
    cir.rylayer(inputs=[0.2 * w for w in range(n)])
    cir.to('cuda', rdtype)
    out = cir().reshape(-1).cpu().numpy().astype(np.complex128)
    ops = [(gates_np.H, [w], []) for w in range(n)] + [(u5, w5, [])]
    ops += [(gates_np.rx(np.float32(0.3 + 0.1 * w)), [w], []) for w in range(n)] + [(gates_np.X, [2], [9]), (u6, w6, [10])]
    ops += [(gates_np.ry(np.float32(0.2 * w)), [w], []) for w in range(n)]
    ref = so.run_circuit(ops, n)
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert err < (1e-6 if rdtype == torch.float64 else 5e-6), err     # float32 parameters (gate.py:384-391)
    # inverse circuit: back to |0...0>
    back = (cir + cir.inverse())().reshape(-1)
    assert abs(abs(complex(back[0])) - 1) < (1e-6 if rdtype == torch.float64 else 1e-4)


@pytest.mark.parametrize('k,adjoint,with_ctrl', [(4, False, False), (5, False, True), (6, False, False), (6, True, True)])
def test_dense_block_on_tensor_cores(k, adjoint, with_ctrl):
    """`b200q_dense_tc_apply` (tcgen05.mma, accumulator in tensor memory, 3-product TF32 split) for dense blocks on 4, 5
    and 6 targets against the oracle in complex128: rel-L2 <= 1e-6 (north_star tolerance; measured 1-4e-7, the FP32
    CUDA-core contraction sits at 1-2e-7)."""
    n = 14
    rng = np.random.default_rng(30 + k)
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi /= np.linalg.norm(psi)
    q, _ = np.linalg.qr(rng.normal(size=(2**k, 2**k)) + 1j * rng.normal(size=(2**k, 2**k)))
    wires = [int(w) for w in rng.permutation(n)[:k]]
    ctrl_wires = [int(w) for w in range(n) if w not in wires][:1] if with_ctrl else []
    u_eff = q.conj().T if adjoint else q
    ref = so.evolve_state_controlled(psi.reshape(1, -1), u_eff, n, wires, ctrl_wires).reshape(-1)
    st = torch.tensor(psi, dtype=torch.complex64, device='cuda')
    u = torch.tensor(q, dtype=torch.complex64, device='cuda').reshape(-1).contiguous()
    t = (C.c_int32 * k)(*[n - 1 - w for w in reversed(wires)])
    ctrl = sum(1 << (n - 1 - c) for c in ctrl_wires)
    L.check(L.load().b200q_dense_tc_apply(st.data_ptr(), n, u.data_ptr(), t, k, ctrl, int(adjoint),
                                          torch.cuda.current_stream().cuda_stream))
    out = st.cpu().numpy().astype(np.complex128)
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert err < 1e-6, err


@pytest.mark.parametrize('rdtype', [torch.float64, torch.float32])
def test_trainable_dense_blocks_gradient_matches_reference_autograd(rdtype):
    """Loss and gradient through trainable dense blocks on 3 and 4 wires (HamiltonianGate in matrix and Pauli-sum
    form, one controlled, and a LatentGate; reference gate.py:2793-3024) against the reference's own autograd
    (tests/golden/dense_grad.npz): such gates get a pass of their own and the reverse sweep accumulates their full
    2^k x 2^k cotangent (b200q_dense_cotangent_kernel)."""
    from test_host_api import _dense_grad_circuit
    g = _golden('dense_grad.npz')
    cir = _dense_grad_circuit(g)
    cir.to('cuda', rdtype)
    data = torch.tensor(g['data'], device='cuda', dtype=rdtype, requires_grad=True)
    state = cir(data)
    tol = 1e-9 if rdtype == torch.float64 else 2e-5
    np.testing.assert_allclose(state.detach().reshape(-1).cpu().numpy(), g['state'], atol=tol)
    exp = cir.expectation()
    np.testing.assert_allclose(exp.detach().reshape(-1).cpu().numpy(), g['expectation'], atol=tol)
    loss = (torch.tensor(g['weights'], device='cuda', dtype=rdtype) * exp.reshape(-1)).sum()
    loss.backward()
    # the float32-rounded Hadamard of the reference bounds the accuracy of ANY adjoint-method gradient at ~1e-6
    gtol = 2e-6 if rdtype == torch.float64 else 2e-4
    np.testing.assert_allclose(data.grad.cpu().numpy(), g['grad/data'], atol=gtol)
    checked = 0
    for key in g.files:
        if key.startswith('grad/') and key != 'grad/data':
            _, i, name = key.split('/')
            got = getattr(cir.operators[int(i)], name).grad
            assert got is not None, key
            np.testing.assert_allclose(got.cpu().numpy(), g[key], atol=gtol, err_msg=key)
            checked += 1
    assert checked == 9


@pytest.mark.parametrize('cdtype', [np.complex128, np.complex64])
def test_reverse_sweep_dense_blocks_cabi(cdtype):
    """b200q_adjoint_run on a 13-qubit plan with trainable dense gates on 3, 4 and 5 wires (one controlled, one stored
    as its adjoint): un-computed state, input cotangent and the full 2^k x 2^k cotangents against PyTorch autograd
    through the CPU port of the reference contraction (the case of tests/test_hostemu.py, here on the device; for
    complex64 the 4- and 5-wire blocks are un-applied by the tensor-core kernel)."""
    import ctypes as C

    from test_hostemu import _dense_grad_case
    n = 13
    ops, psi0, psi_f, lam_f, x0_grad, mgrads = _dense_grad_case(n, np.random.default_rng(12))
    need = [1 if (len(e) > 3 and e[3].get('grad')) or len(e[1]) == 1 else 0 for e in ops]
    need = [0 if np.array_equal(np.asarray(e[0]), gates_np.X) else v for e, v in zip(ops, need)]
    arr, ng, mats = lower_ops(ops, n, cdtype)
    tdt = torch.complex64 if cdtype == np.complex64 else torch.complex128
    plan = engine.FusedPlan(n, tdt, list(arr)[:ng], chunk_bits=11)
    psi = torch.tensor(psi_f, dtype=tdt, device='cuda').contiguous()
    lam = torch.tensor(lam_f, dtype=tdt, device='cuda').contiguous()
    m = torch.tensor(mats, device='cuda')
    grad = torch.zeros(m.numel(), dtype=torch.complex128, device='cuda')
    need_c = (C.c_uint8 * len(need))(*need)
    lib = L.load()
    L.check(lib.b200q_adjoint_run(plan._h, psi.data_ptr(), lam.data_ptr(), m.data_ptr(), grad.data_ptr(), need_c,
                                  engine._stream(psi)))
    torch.cuda.synchronize()
    tol = 1e-10 if cdtype == np.complex128 else 3e-4
    assert np.linalg.norm(psi.cpu().numpy() - psi0) < (1e-10 if cdtype == np.complex128 else 1e-4)
    assert np.linalg.norm(lam.cpu().numpy() - x0_grad) / np.linalg.norm(x0_grad) < tol
    grad = grad.cpu().numpy()
    off, checked = 0, 0
    for e, ref, nd in zip(ops, mgrads, need):
        mm = np.asarray(e[0])
        if nd and len(e[1]) >= 3:
            gq = grad[off:off + mm.size].reshape(mm.shape)
            assert np.abs(gq - ref).max() / max(1.0, np.abs(ref).max()) < tol, (e[1], e[2])
            checked += 1
        off += mm.size
    assert checked == 4
