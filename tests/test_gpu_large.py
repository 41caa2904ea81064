"""Amplitude-level parity at the sizes the bench measures ("same generator at
n in {20, 24, 26}"): the C2 generator (`workloads.random_clifford_rx_spec`) at 20, 22, 24 qubits depth 40
and 26 qubits depth 6, full final state of the CUDA path against `oracle/torch_port.run_ops` in
complex128 on the host cores.  At these sizes every pass has several groups of non-tile bits
(n >= 19), the regime of the headline bench.

Run on the B200 box: pytest -m gpu."""
import functools

import numpy as np
import pytest
import torch

import gates_np
import torch_port

import deepquantum_b200 as dq
from deepquantum_b200 import workloads as wl

pytestmark = pytest.mark.gpu

# complex128 engine vs complex128 oracle: rel-L2 <= 1e-10.  complex64 engine vs complex128 oracle:
# rel-L2 <= 2e-6 at depth <= 40 and per-amplitude |delta| <= 1e-6 * max|amp| ... the reference's own
# complex64 path sits at 6-8e-7 rel-L2 at depth 40.
REL_L2 = {torch.float64: 1e-10, torch.float32: 2e-6}
CASES = [(20, 40), (22, 40), (24, 40), (26, 6)]


@functools.lru_cache(maxsize=None)
def _oracle_state(n, depth):
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ops = gates_np.lower_spec(wl.random_clifford_rx_spec(n, depth), n)
    out, done, _ = torch_port.run_ops(ops, n, dtype=torch.complex128)
    assert done == len(ops)
    return out.numpy()


@pytest.mark.parametrize('n,depth', CASES)
@pytest.mark.parametrize('rdtype', [torch.float32, torch.float64])
def test_c2_generator_full_state(n, depth, rdtype):
    ref = _oracle_state(n, depth)
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, wl.random_clifford_rx_spec(n, depth))
    cir.to('cuda', rdtype)
    out = cir().reshape(-1)
    got = out.cpu().numpy().astype(np.complex128)
    del out
    plan = cir._get_program().plan(cir.state.dtype)
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert err < REL_L2[rdtype], (n, depth, rdtype, err, plan.n_passes)
    if rdtype == torch.float32:
        amax = np.abs(ref).max()
        # per-amplitude bound (SURVEY 8d states 1e-6 * max|amp|; 2e-6 leaves room for the float32 tail at depth 40)
        assert np.abs(got - ref).max() <= 2e-6 * amax, np.abs(got - ref).max() / amax


def test_qaoa_c3_shape_loss_and_gradient_n16():
    """Config 3 shape at 16 qubits (p = 2), complex128: loss and its gradient (adjoint sweep kernel) against dense torch
    autograd through the oracle's gate loop on the host."""
    n, p = 16, 2
    edges, weights = wl.random_regular_graph(n, 3)
    params = torch.tensor([0.1] * p + [1.0] * p, dtype=torch.float64, requires_grad=True)

    def build(device):
        cir = dq.QubitCircuit(n)
        cir.hlayer()
        for _ in range(p):
            for (a, b) in edges:
                cir.cnot(a, b)
                cir.rz(b, encode=True)
                cir.cnot(a, b)
            for i in range(n):
                cir.rx(i, encode=True)
        for (a, b) in edges:
            cir.observable([a, b], 'zz')
        cir.to(device, torch.double)
        return cir

    def expand(prm):
        w = torch.tensor(weights, dtype=torch.float64, device=prm.device)
        parts = []
        for k in range(p):
            parts.append(prm[k] * w)
            parts.append(prm[p + k].expand(n))
        return torch.cat(parts)

    cir = build('cuda')
    pg = params.detach().clone().cuda().requires_grad_(True)
    cir(expand(pg))
    loss = (cir.expectation().reshape(-1) * torch.tensor(weights, dtype=torch.float64, device='cuda')).sum()
    loss.backward()

    # host: dense autograd with torch (complex128), gate by gate like the reference
    def rz_m(t):
        e = torch.exp(-0.5j * t.to(torch.complex128))
        return torch.stack([e, torch.zeros_like(e), torch.zeros_like(e), e.conj()]).reshape(2, 2)

    def rx_m(t):
        c, s = torch.cos(t / 2).to(torch.complex128), torch.sin(t / 2).to(torch.complex128)
        return torch.stack([c, -1j * s, -1j * s, c]).reshape(2, 2)

    ph = params.detach().clone().requires_grad_(True)
    data = expand(ph)
    hm = torch.tensor(gates_np.H, dtype=torch.complex128)
    xm = torch.tensor(gates_np.X, dtype=torch.complex128)
    st = torch.zeros(2**n, dtype=torch.complex128)
    st[0] = 1
    x = st.reshape([1] + [2] * n)
    for i in range(n):
        x = torch_port.evolve_state(x, hm, n, [i])
    j = 0
    for _ in range(p):
        for (a, b) in edges:
            x = torch_port.evolve_state_controlled(x, xm, n, [b], [a])
            x = torch_port.evolve_state(x, rz_m(data[j]), n, [b])
            j += 1
            x = torch_port.evolve_state_controlled(x, xm, n, [b], [a])
        for i in range(n):
            x = torch_port.evolve_state(x, rx_m(data[j]), n, [i])
            j += 1
    psi = x.reshape(-1)
    prob = (psi.conj() * psi).real
    idx = torch.arange(2**n)
    loss_ref = 0
    for (a, b), w in zip(edges, weights):
        sign = 1 - 2 * (((idx >> (n - 1 - a)) ^ (idx >> (n - 1 - b))) & 1).to(torch.float64)
        loss_ref = loss_ref + w * (prob * sign).sum()
    loss_ref.backward()
    assert abs(float(loss) - float(loss_ref)) < 1e-6 * max(1.0, abs(float(loss_ref)))
    g, gr = pg.grad.cpu().numpy(), ph.grad.numpy()
    assert np.abs(g - gr).max() < 2e-6 * max(1.0, np.abs(gr).max()), (g, gr)


@pytest.mark.parametrize('world', [2, 8])
@pytest.mark.parametrize('rdtype', [torch.float64, torch.float32])
def test_sharded_schedule_all_ranks_on_one_gpu(world, rdtype):
    """The sharded path on a ONE-GPU box: every rank's `ShardedProgram` ('perm' schedule: exchanges are bit
    permutations fused into the last pass of a segment) is stepped in lockstep in this process, every rank's shard and
    receive buffer are tensors on cuda:0, and the "peer" pointers of the fused exchange are simply the other ranks'
    buffers -- the exchange kernel, the rank predicates / per-rank phases and the specialised pass kernels (shards of
    >= 2^20 amplitudes) all run on the GPU.  C2 generator at 23 qubits depth 12 against the CPU oracle."""
    from deepquantum_b200.distributed import CudaExecutor, ShardedProgram
    n, depth = 23, 12
    g = world.bit_length() - 1
    nl = n - g
    spec = wl.random_clifford_rx_spec(n, depth)
    ops = gates_np.lower_spec(spec, n)
    ref, done, _ = torch_port.run_ops(ops, n, dtype=torch.complex128)
    ref = ref.numpy()
    dense = dq.QubitCircuit(n)
    wl.apply_spec(dense, spec)
    dense.to('cuda', rdtype)
    cdt = torch.complex128 if rdtype == torch.float64 else torch.complex64
    low = dense._get_program().low
    mats = low.build_matrices(cdt, 'cuda').detach()

    class State:
        def __init__(self, r):
            self.amps = torch.zeros(2**nl, dtype=cdt, device='cuda')
            self.buffer = torch.zeros(2**nl, dtype=cdt, device='cuda')
            if r == 0:
                self.amps[0] = 1.0

        def enable_peer_exchange(self):
            return True

        def peer_buffer_ptrs(self):
            return [s.buffer.data_ptr() for s in states]

    states = [State(r) for r in range(world)]
    progs = [ShardedProgram(low, n, world, r, 'perm') for r in range(world)]
    ex = CudaExecutor()
    for p in progs:
        p.fused_exchanges, p._skip_next = 0, False
    assert any(s[0] == 'xperm' for s in progs[0].steps)
    for si in range(len(progs[0].steps)):
        what = [progs[r].run_step(si, states[r], mats, ex, True) for r in range(world)]
        assert len(set(what)) == 1, what
        torch.cuda.synchronize()
        if what[0] == 'exchange':
            for r in range(world):
                ShardedProgram.commit_exchange(states[r])
    got = torch.cat([s.amps for s in states]).cpu().numpy().astype(np.complex128)
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert err < REL_L2[rdtype], (world, rdtype, err)
    assert progs[0].fused_exchanges > 0


def _remap(spec, wires_of):
    """The spec of a k-qubit circuit moved onto the wires `wires_of[0..k-1]` of a larger register."""
    out = []
    for e in spec:
        f = dict(e)
        for key in ('w', 'c'):
            if key in f:
                f[key] = [wires_of[q] for q in f[key]]
        out.append(f)
    return out


@pytest.mark.parametrize('n,na,rdtype', [(30, 20, torch.float32), (29, 20, torch.float64)])
def test_headline_size_full_state_against_the_oracle_by_tensor_product(n, na, rdtype):
    """Every amplitude of a 30-QUBIT run (the metric's size: 8 GiB, 17 non-tile bits, the same specialised kernels as the
    bench) against the oracle: the C2 generator at depth 40 runs on 20 of the 30 wires and, interleaved gate by gate, at
    depth 40 on the other 10 -- the two wire sets are scattered over high and low index bits -- so the exact final state is
    the tensor product of a 20-qubit and a 10-qubit oracle state (complex128, host); complex128 at 29 qubits (8 GiB too).  Size-independent property used:
    gates on disjoint wire sets factorise; nothing about the engine's passes does (both circuits share every pass)."""
    nb, depth = n - na, 40
    rng = np.random.default_rng(30)
    perm = rng.permutation(n)
    wires_a, wires_b = sorted(perm[:na].tolist()), sorted(perm[na:].tolist())
    spec_a, spec_b = wl.random_clifford_rx_spec(na, depth), wl.random_clifford_rx_spec(nb, depth, seed=wl.SEED + 1)
    ref_a = _oracle_state(na, depth)
    ops_b = gates_np.lower_spec(spec_b, nb)
    ref_b, done, _ = torch_port.run_ops(ops_b, nb, dtype=torch.complex128)
    assert done == len(ops_b)
    ga, gb = _remap(spec_a, wires_a), _remap(spec_b, wires_b)
    merged, ia, ib = [], 0, 0
    while ia < len(ga) or ib < len(gb):        # interleave 2 : 1, keeping each circuit's own order
        for _ in range(2):
            if ia < len(ga):
                merged.append(ga[ia])
                ia += 1
        if ib < len(gb):
            merged.append(gb[ib])
            ib += 1
    cir = dq.QubitCircuit(n)
    wl.apply_spec(cir, merged)
    cir.to('cuda', rdtype)
    out = cir().reshape(-1)
    plan = cir._get_program().plan(out.dtype)
    assert plan.jit_status()['specialised'] == plan.n_passes
    ta = torch.tensor(ref_a, device='cuda')
    tb = ref_b.reshape(-1).to('cuda')
    pos_a = [n - 1 - w for w in wires_a]       # index bit of wire w; wire order = most significant first
    pos_b = [n - 1 - w for w in wires_b]
    err2 = torch.zeros((), dtype=torch.float64, device='cuda')
    worst = torch.zeros((), dtype=torch.float64, device='cuda')
    chunk = 1 << 24
    for start in range(0, 1 << n, chunk):
        idx = torch.arange(start, start + chunk, device='cuda', dtype=torch.int64)
        ia_ = torch.zeros_like(idx)
        for k, p in enumerate(pos_a):
            ia_ |= ((idx >> p) & 1) << (na - 1 - k)
        ib_ = torch.zeros_like(idx)
        for k, p in enumerate(pos_b):
            ib_ |= ((idx >> p) & 1) << (nb - 1 - k)
        want = ta[ia_] * tb[ib_]
        diff = (out[start:start + chunk].to(torch.complex128) - want).abs()
        err2 += (diff**2).sum()
        worst = torch.maximum(worst, diff.max())
    rel = float(err2.sqrt())                    # the product state has norm 1
    amax = float(ta.abs().max() * tb.abs().max())
    tol = REL_L2[rdtype]
    assert rel < tol, rel
    assert float(worst) <= tol * amax, float(worst) / amax
