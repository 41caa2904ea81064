"""Sampling (qmath.measure, reference qmath.py:543-638): the oracle against the reference-generated fixture
(CPU), the CUDA kernels against the oracle (GPU, through the C ABI and through the product API)."""
import json
import os

import numpy as np
import pytest

import sampling_oracle as smp
from conftest import GOLDEN


def _fixture():
    return np.load(os.path.join(GOLDEN, 'measure.npz'))


CASES = ['all5', 'sub6', 'one7', 'batch5']


@pytest.mark.parametrize('case', CASES)
def test_oracle_probabilities_match_reference(case):
    """The probabilities (and the key convention: sorted wires, wire 0 first) the unmodified reference attaches
    to its outcomes are reproduced by the oracle's marginal distribution."""
    g = _fixture()
    meta = json.loads(str(g[case + '/meta']))
    n, wires = meta['n'], meta['wires']
    for b in range(meta['batch']):
        p = smp.probabilities(g[case + '/state'][b], n, wires)
        keys = [str(k) for k in g[f'{case}/keys{b}']]
        ref = g[f'{case}/probs{b}']
        got = np.array([p[int(k, 2)] for k in keys])
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-15)
        nbits = n if wires is None else (1 if isinstance(wires, int) else len(wires))
        assert all(len(k) == nbits for k in keys)
        # the reference's counts are a plausible draw from the oracle's distribution (4096 shots)
        counts = g[f'{case}/counts{b}']
        exp = got * counts.sum()
        big = exp > 5
        chi2 = float(((counts[big] - exp[big])**2 / exp[big]).sum())
        assert chi2 < 3.0 * max(1, big.sum()), (chi2, int(big.sum()))


def test_oracle_inverse_cdf_known_answers():
    n = 6
    basis = np.zeros(2**n, complex)
    basis[37] = 1.0
    assert set(smp.sample_indices(basis, np.linspace(0, 0.999, 50))) == {37}
    uni = np.ones(2**n, complex) / np.sqrt(2**n)
    u = (np.arange(2**n) + 0.5) / 2**n
    assert np.array_equal(smp.sample_indices(uni, u), np.arange(2**n))
    assert smp.measure(basis, n, np.array([0.1, 0.7]), wires=[5, 0, 2]) == {'101': 2}   # 37 = 100101b
    ghz = np.zeros(2**n, complex)
    ghz[0] = ghz[-1] = np.sqrt(0.5)
    assert smp.measure(ghz, n, np.array([0.25, 0.75, 0.9])) == {'0' * n: 1, '1' * n: 2}


# ---------------------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


def _rand_state(n, seed, cdtype):
    rng = np.random.default_rng(seed)
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    psi *= rng.uniform(0, 1, size=2**n) ** 4       # uneven masses, some nearly empty blocks
    return (psi / np.linalg.norm(psi)).astype(cdtype)


@gpu
@pytest.mark.parametrize('cdtype', [np.complex64, np.complex128])
@pytest.mark.parametrize('n', [3, 11, 14, 20])
def test_sample_indices_match_oracle(n, cdtype):
    """Same uniforms -> same indices as the oracle's inverse CDF, up to the summation order at a CDF boundary:
    the kernel's index must bracket u within 1e-12 of the total mass."""
    import torch
    from deepquantum_b200 import engine
    psi = _rand_state(n, n, cdtype)
    u = np.random.default_rng(1).uniform(size=2000)
    st = torch.from_numpy(psi).cuda()
    idx = engine.sample_indices(st, n, torch.from_numpy(u)).cpu().numpy()
    ref = smp.sample_indices(psi, u)
    p = np.abs(psi.astype(np.complex128))**2
    cdf = np.cumsum(p)
    lo = np.where(idx > 0, cdf[np.maximum(idx, 1) - 1], 0.0)
    hi = cdf[idx]
    t = u * cdf[-1]
    assert np.all(p[idx] > 0)
    assert np.all(lo - 1e-12 <= t) and np.all(t <= hi + 1e-12)
    assert np.mean(idx == ref) > 0.999


@gpu
def test_block_mass_and_marginal_probs():
    import torch
    from deepquantum_b200 import engine
    n = 16
    psi = _rand_state(n, 5, np.complex64)
    st = torch.from_numpy(psi).cuda()
    p = np.abs(psi.astype(np.complex128))**2
    for bb in (0, 1, 7, 12, 16):
        m = engine.block_mass(st, n, 1, bb)[0].cpu().numpy()
        np.testing.assert_allclose(m, p.reshape(-1, 2**bb).sum(1), rtol=1e-12, atol=1e-18)
    wires = [2, 9, 15, 4]
    marg = smp.probabilities(psi, n, wires)
    mask = sum(1 << (n - 1 - w) for w in wires)
    sw = sorted(wires)
    keys = np.arange(2**len(wires))
    dep = np.zeros_like(keys)
    for j, w in enumerate(sw):
        dep |= ((keys >> (len(sw) - 1 - j)) & 1) << (n - 1 - w)
    order = np.argsort(dep)
    got = engine.marginal_probs(st, n, int(mask), torch.from_numpy(dep[order].astype(np.int64)).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, marg[order], rtol=1e-10)


@gpu
@pytest.mark.parametrize('case', CASES)
def test_measure_api_against_fixture(case):
    """dq.measure / the reference fixture: identical keys convention and probabilities; counts are a plausible
    draw (chi-square against the exact distribution)."""
    import torch
    import deepquantum_b200 as dq
    from deepquantum_b200 import qmath
    g = _fixture()
    meta = json.loads(str(g[case + '/meta']))
    n, wires, batch = meta['n'], meta['wires'], meta['batch']
    st = torch.from_numpy(g[case + '/state']).cuda()
    arg = st[0] if batch == 1 else st
    gen = torch.Generator().manual_seed(7)
    res = qmath.measure(arg, shots=4096, with_prob=True, wires=wires, generator=gen)
    res = [res] if batch == 1 else res
    assert len(res) == batch
    for b, d in enumerate(res):
        p = smp.probabilities(g[case + '/state'][b], n, wires)
        tot = 0
        chi2, dof = 0.0, 0
        for k, (c, pr) in d.items():
            assert abs(float(pr) - p[int(k, 2)]) < 1e-12
            tot += c
        assert tot == 4096
        for i, pi in enumerate(p):
            e = pi * 4096
            if e > 5:
                c = d.get(format(i, f'0{len(next(iter(d)))}b'), (0, 0))[0]
                chi2 += (c - e)**2 / e
                dof += 1
        assert chi2 < 3.0 * max(1, dof), (chi2, dof)
        ref_keys = {str(k) for k in g[f'{case}/keys{b}']}
        assert len(next(iter(d))) == len(next(iter(ref_keys)))


@gpu
def test_circuit_measure_full_size_properties():
    """24-qubit GHZ-like circuit: only the two GHZ strings are ever drawn; marginal on 3 wires likewise."""
    import deepquantum_b200 as dq
    n = 24
    cir = dq.QubitCircuit(n)
    cir.h(0)
    for i in range(n - 1):
        cir.cnot(i, i + 1)
    cir.to('cuda:0')
    cir()
    r = cir.measure(shots=2000)
    assert set(r) <= {'0' * n, '1' * n} and sum(r.values()) == 2000
    assert abs(r.get('0' * n, 0) - 1000) < 150
    r = cir.measure(shots=500, wires=[3, 20, 11], with_prob=True)
    assert set(r) <= {'000', '111'}
    for k, (c, p) in r.items():
        assert abs(float(p) - 0.5) < 1e-6
