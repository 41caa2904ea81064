"""Seeded synthetic workloads of BASELINE.json / SURVEY.md section 8(d), as circuit *specs*.

A spec is a JSON-able list of ``{"g": builder-name, "w": wires, "c": controls, "p": params}``
entries; `apply_spec` replays it through the `QubitCircuit` builder API (the reference's method
names, circuit.py:899-1537), so the same spec drives the product engine, the reference (golden
generation) and the CPU oracle.
"""
from __future__ import annotations

import math

import torch

SEED = 20261017


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def c1_plumbing_spec(nqubit: int = 12, seed: int = SEED):
    """Config 1: hlayer, CNOT ladder, rxlayer (35 gates at 12 qubits)."""
    g = _gen(seed)
    theta = (torch.rand(nqubit, generator=g, dtype=torch.float64) * 2 * math.pi).tolist()
    spec = [{'g': 'hlayer'}]
    spec += [{'g': 'cnot', 'w': [i, i + 1]} for i in range(nqubit - 1)]
    spec.append({'g': 'rxlayer', 'p': theta})
    return spec


def random_clifford_rx_spec(nqubit: int, depth: int, seed: int = SEED, two_qubit: str = 'cnot'):
    """Configs 2 and 4: per layer one gate from {H, S, RX(theta)} on every qubit, then CNOTs on a
    random perfect matching of a random qubit permutation (arbitrary control/target distance)."""
    g = _gen(seed)
    spec = []
    for _ in range(depth):
        kinds = torch.randint(0, 3, (nqubit,), generator=g).tolist()
        angles = (torch.rand(nqubit, generator=g, dtype=torch.float64) * 2 * math.pi).tolist()
        for q in range(nqubit):
            if kinds[q] == 0:
                spec.append({'g': 'h', 'w': [q]})
            elif kinds[q] == 1:
                spec.append({'g': 's', 'w': [q]})
            else:
                spec.append({'g': 'rx', 'w': [q], 'p': [angles[q]]})
        perm = torch.randperm(nqubit, generator=g).tolist()
        for i in range(0, nqubit - 1, 2):
            spec.append({'g': two_qubit, 'w': [perm[i], perm[i + 1]]})
    return spec


def random_regular_graph(nnodes: int, degree: int = 3, seed: int = SEED):
    """Seeded random d-regular simple graph (pairing model with restarts) + U[0.3,0.9] weights
    (the `examples/qaoa.py:89-99` shape)."""
    g = _gen(seed)
    assert nnodes * degree % 2 == 0
    while True:
        stubs = [v for v in range(nnodes) for _ in range(degree)]
        order = torch.randperm(len(stubs), generator=g).tolist()
        stubs = [stubs[i] for i in order]
        edges = set()
        ok = True
        for i in range(0, len(stubs), 2):
            a, b = stubs[i], stubs[i + 1]
            if a == b or (min(a, b), max(a, b)) in edges:
                ok = False
                break
            edges.add((min(a, b), max(a, b)))
        if ok:
            break
    edges = sorted(edges)
    weights = (torch.rand(len(edges), generator=g, dtype=torch.float64) * 0.6 + 0.3).tolist()
    return edges, weights


def qaoa_maxcut_structure(nqubit: int, p: int = 4, seed: int = SEED):
    """Config 3 (examples/qaoa.py:31-53): hlayer; per step, for each edge cnot-rz-cnot (angle
    gamma_k * w_e) and rx(beta_k) on every qubit; observables ZZ per edge.

    Returns (edges, weights, layout) where layout lists, per encoded angle in circuit order,
    ('gamma', k, edge_index) or ('beta', k, qubit)."""
    edges, weights = random_regular_graph(nqubit, 3, seed)
    layout = []
    for k in range(p):
        for ei in range(len(edges)):
            layout.append(('gamma', k, ei))
        for q in range(nqubit):
            layout.append(('beta', k, q))
    return edges, weights, layout


def build_qaoa(cir, edges, p: int, barriers: bool = True):
    """Add the QAOA gates (all angles encoded, data supplied at call time) and ZZ observables."""
    n = cir.nqubit
    cir.hlayer()
    for _ in range(p):
        for a, b in edges:
            cir.cnot(a, b)
            cir.rz(b, encode=True)
            cir.cnot(a, b)
        if barriers:
            cir.barrier()
        for q in range(n):
            cir.rx(q, encode=True)
        if barriers:
            cir.barrier()
    for a, b in edges:
        cir.observable([a, b], 'z')
    return cir


def qaoa_data(params: torch.Tensor, weights, layout):
    """Expand the 2p scalar parameters (gamma_0..gamma_{p-1}, beta_0..beta_{p-1}) into the encoded
    angle vector (examples/qaoa.py:46-53): rz angle = 2*gamma_k*w_e, rx angle = 2*beta_k."""
    p = params.numel() // 2
    w = torch.as_tensor(weights, dtype=params.dtype, device=params.device)
    out = []
    for kind, k, idx in layout:
        if kind == 'gamma':
            out.append(2 * params[k] * w[idx])
        else:
            out.append(2 * params[p + k])
    return torch.stack(out)


def fock_interferometer_spec(nmode: int = 8, seed: int = SEED):
    """Config 5: squeezer on every mode, then a rectangular (Clements) mesh of beamsplitters."""
    g = _gen(seed)
    r = (torch.rand(nmode, generator=g, dtype=torch.float64) * 0.5).tolist()
    th = (torch.rand(nmode, generator=g, dtype=torch.float64) * 2 * math.pi).tolist()
    spec = [{'g': 's', 'w': [i], 'p': [r[i], th[i]]} for i in range(nmode)]
    for layer in range(nmode):
        for i in range(layer % 2, nmode - 1, 2):
            a = (torch.rand(2, generator=g, dtype=torch.float64) * 2 * math.pi).tolist()
            spec.append({'g': 'bs', 'w': [i, i + 1], 'p': a})
    return spec


def noisy_circuit_spec(nqubit: int, depth: int, seed: int = SEED, channels=None):
    """Density-matrix workload: the Clifford+RX layers of `random_clifford_rx_spec`
    with one channel per qubit after every layer, cycling through the seven channel families (or `channels`)."""
    g = _gen(seed)
    base = random_clifford_rx_spec(nqubit, depth, seed)
    per_layer = len(base) // depth
    names = list(channels) if channels else ['bit_flip', 'phase_flip', 'depolarizing', 'pauli', 'amp_damp',
                                             'phase_damp', 'gen_amp_damp']
    spec, k = [], 0
    for d in range(depth):
        spec += base[d * per_layer:(d + 1) * per_layer]
        for q in range(nqubit):
            name = names[k % len(names)]
            k += 1
            npar = {'pauli': 4, 'gen_amp_damp': 2}.get(name, 1)
            prm = (torch.rand(npar, generator=g, dtype=torch.float64) * 0.6 + 0.05).tolist()
            spec.append({'g': name, 'w': [q], 'p': prm})
    return spec


def count_gates(spec, nqubit):
    n = 0
    for e in spec:
        g = e['g']
        if g.endswith('layer'):
            n += len(e.get('w') or range(nqubit)) if g != 'cxlayer' else len(e['pairs'])
        elif g == 'cnot_ring':
            lo, hi = e.get('minmax') or [0, nqubit - 1]
            n += hi - lo + 1
        elif g != 'barrier':
            n += 1
    return n


_CONST_1Q = ('x', 'y', 'z', 'h', 's', 'sdg', 't', 'tdg')
_PARAM_2Q = ('rxx', 'ryy', 'rzz', 'rxy', 'rbs')


def apply_spec(cir, spec, cdtype=torch.complex64):
    """Replay a spec through the builder API shared by the reference `QubitCircuit` and the product
    `deepquantum_b200.QubitCircuit` (same method names and argument meaning)."""

    for e in spec:
        g = e['g']
        w = list(e.get('w', []))
        c = list(e.get('c', [])) or None
        prm = e.get('p', [])
        if g in _CONST_1Q:
            getattr(cir, g)(w[0], controls=c)
        elif g in ('rx', 'ry', 'rz', 'p'):
            getattr(cir, g)(w[0], prm[0], controls=c)
        elif g == 'u3':
            cir.u3(w[0], list(prm), controls=c)
        elif g == 'j':
            cir.j(w[0], prm[0], plane=e.get('plane', 'xy'), controls=c)
        elif g in ('cx', 'cy', 'cz', 'ch', 'cs', 'csdg', 'ct', 'ctdg', 'cnot'):
            getattr(cir, g)(w[0], w[1])
        elif g in ('crx', 'cry', 'crz', 'cp'):
            getattr(cir, g)(w[0], w[1], prm[0])
        elif g == 'cu':
            cir.cu(w[0], w[1], list(prm))
        elif g in ('swap', 'iswap'):
            getattr(cir, g)(w, controls=c)
        elif g in _PARAM_2Q:
            getattr(cir, g)(w, prm[0], controls=c)
        elif g in ('crxx', 'cryy', 'crzz', 'crxy'):
            getattr(cir, g)(w[0], w[1], w[2], prm[0])
        elif g in ('toffoli', 'ccx', 'fredkin', 'cswap'):
            getattr(cir, g)(w[0], w[1], w[2])
        elif g == 'any':
            u = torch.complex(torch.tensor(e['u_re'], dtype=torch.float64), torch.tensor(e['u_im'], dtype=torch.float64))
            cir.any(u.to(cdtype), wires=w, controls=c)
        elif g in ('xlayer', 'ylayer', 'zlayer', 'hlayer'):
            getattr(cir, g)(w or None)
        elif g in ('rxlayer', 'rylayer', 'rzlayer', 'u3layer'):
            getattr(cir, g)(w or None, list(prm))
        elif g == 'cxlayer':
            cir.cxlayer([list(pr) for pr in e['pairs']])
        elif g == 'cnot_ring':
            cir.cnot_ring(e.get('minmax'), e.get('step', 1), e.get('reverse', False))
        elif g == 'barrier':
            cir.barrier()
        elif g == 'hamiltonian':
            if 'ham' in e:       # Pauli-sum list, spans the min..max wire it names
                cir.hamiltonian(e['ham'], prm[0], controls=c)
            else:
                h = torch.complex(torch.tensor(e['h_re'], dtype=torch.float64),
                                  torch.tensor(e['h_im'], dtype=torch.float64))
                cir.hamiltonian(h.to(cdtype), prm[0], wires=w, controls=c)
        elif g in ('bit_flip', 'phase_flip', 'depolarizing', 'amp_damp', 'phase_damp'):   # den_mat circuits only
            getattr(cir, g)(w[0], prm[0])
        elif g in ('pauli', 'gen_amp_damp'):
            getattr(cir, g)(w[0], list(prm))
        else:
            raise ValueError(g)
    return cir
