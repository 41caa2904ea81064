"""Sharded statevector: the high-order qubit index is split over the ranks (reference state.py:342-383,
distributed.py:57-202, circuit.py:1625-1770), re-designed around ONE rule: the local fused engine only ever
sees local qubits.

  * rank r holds flat indices [r * 2^(n-g), (r+1) * 2^(n-g)), g = log2(world size) (wires 0..g-1 global);
  * a persistent logical -> physical qubit map replaces the reference's swap-in / apply / swap-out
    (distributed.py:194-201): gates run fused on the local shard while their non-diagonal targets are
    local; controls and diagonal gates on global qubits need NO data exchange (they become per-rank
    predicates / phases, distributed.py:84-93);
  * when the front of the circuit is blocked by global targets, ONE all-to-all block transpose swaps all g
    global qubits with the top g local qubits (each rank sends (W-1)/W of its shard once);
  * the map is restored before the state is handed back, so `.amps` has the reference layout.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import _lib as L
from . import engine
from .communication import block_transpose, comm_get_rank, comm_get_world_size
from .operation import Lowering, apply_complex_fix


class DistributedQubitState(nn.Module):
    """Per-rank shard `amps` plus a same-size `buffer` (reference state.py:342-383)."""

    def __init__(self, nqubit: int) -> None:
        super().__init__()
        self.world_size = comm_get_world_size()
        self.rank = comm_get_rank()
        assert self.world_size & (self.world_size - 1) == 0
        assert 2**nqubit >= self.world_size
        self.nqubit = nqubit
        self.log_num_nodes = self.world_size.bit_length() - 1
        self.log_num_amps_per_node = nqubit - self.log_num_nodes
        self.num_amps_per_node = 2**self.log_num_amps_per_node
        self.register_buffer('amps', torch.zeros(self.num_amps_per_node) + 0j)
        self.register_buffer('buffer', torch.zeros(self.num_amps_per_node) + 0j)
        self.reset()

    def _apply(self, fn):
        tensors = {k: self._buffers.pop(k) for k in ('amps', 'buffer')}
        super()._apply(fn)
        for key, value in apply_complex_fix(fn, tensors).items():
            self.register_buffer(key, value)
        return self

    def reset(self) -> None:
        self.amps.zero_()
        if self.rank == 0:
            self.amps[0] = 1.0

    # ---- peer mappings for the fused pass + exchange -----------------------------------------------------------
    def enable_peer_exchange(self) -> bool:
        """Re-home `amps` and `buffer` in symmetric memory (torch.distributed._symmetric_memory: one allocation per
        rank, mapped into every other rank of the node over NVLink) so that the last pass of a local segment can
        store straight into the peers' receive buffers.  Collective; returns False (and keeps the NCCL transpose)
        when symmetric memory is not available.  `B200Q_PEER_EXCHANGE=0` disables it."""
        import os

        import torch.distributed as dist
        peer = self.__dict__.get('_peer')
        if isinstance(peer, dict) and self.amps.data_ptr() in peer and self.buffer.data_ptr() in peer:
            return True
        if peer is False and self.__dict__.get('_peer_key') == (self.amps.dtype, str(self.amps.device)):
            return False
        self.__dict__['_peer_key'] = (self.amps.dtype, str(self.amps.device))
        ok = (self.world_size > 1 and self.world_size <= 8 and self.amps.is_cuda and dist.is_initialized()
              and dist.get_backend() == 'nccl' and os.environ.get('B200Q_PEER_EXCHANGE', '1') != '0')
        err = ''
        if ok:
            try:
                import torch.distributed._symmetric_memory as symm
                group = dist.group.WORLD
                a = symm.empty(self.num_amps_per_node, dtype=self.amps.dtype, device=self.amps.device)
                b = symm.empty(self.num_amps_per_node, dtype=self.amps.dtype, device=self.amps.device)
                ha = symm.rendezvous(a, group)
                hb = symm.rendezvous(b, group)
                a.copy_(self.amps)
                b.zero_()
                peers = {a.data_ptr(): [int(x) for x in ha.buffer_ptrs], b.data_ptr(): [int(x) for x in hb.buffer_ptrs]}
                assert len(peers[a.data_ptr()]) == self.world_size
            except Exception as e:       # noqa: BLE001 -- any failure keeps the NCCL path
                ok, err = False, f'{type(e).__name__}: {e}'
        # every rank must take the same path
        flag = torch.tensor([1 if ok else 0], device=self.amps.device if self.amps.is_cuda else 'cpu')
        if dist.is_initialized() and self.world_size > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if not bool(flag.item()):
            self.__dict__['_peer'] = False
            self.__dict__['_peer_error'] = err or 'disabled or unavailable on some rank'
            return False
        self.register_buffer('amps', a)
        self.register_buffer('buffer', b)
        self.__dict__['_peer'] = peers
        self.__dict__['_peer_handles'] = (ha, hb)     # keep the mappings alive
        return True

    def peer_buffer_ptrs(self):
        """Device pointers, valid on this device, of every rank's CURRENT receive buffer (all ranks swap the roles
        of `amps` and `buffer` at the same steps, so the parity is the same everywhere)."""
        return self.__dict__['_peer'][self.buffer.data_ptr()]


class CudaExecutor:
    """Runs one local segment (a fused plan over the shard) on the GPU through the C ABI."""

    def make_plan(self, nlocal, dtype, structs, exchange=False):
        # a segment whose last pass stores over NVLink writes 128-byte runs per quarter warp (3 lane-owned chunk bits)
        return engine.FusedPlan(nlocal, dtype, structs, coalesce_bits=3 if exchange else None)

    def run_plan(self, plan, amps, mats):
        plan.run(amps, mats, 1, 0)

    def run_plan_exchange(self, plan, amps, mats, peer_ptrs, rank, perm=None):
        """The segment's plan with its last pass storing into the peers' receive buffers (b200q_plan_run_exchange);
        `perm`: bit permutation of the distributed index, None = block transpose."""
        plan.run_exchange(amps, mats, peer_ptrs, rank, perm)

    # local pieces of the (differentiable) expectation: fused Z-string reduction, its cotangent, the reverse sweep
    def expectation_z(self, amps, nlocal, masks, index_offset):
        return engine.expectation_z(amps, nlocal, masks, 1, index_offset=index_offset).reshape(-1)

    def apply_z_weights(self, amps, nlocal, masks, weights, index_offset):
        return engine.apply_z_weights(amps, nlocal, masks, weights.reshape(1, -1), 1, index_offset=index_offset)

    def run_plan_adjoint(self, plan, psi, lam, mats, grad, need):
        """psi <- U^dagger psi, lam <- U^dagger lam, grad += cotangent of the matrices (b200q_adjoint_run)."""
        import ctypes as C
        arr = (C.c_uint8 * max(1, len(need)))(*need)
        L.check(L.load().b200q_adjoint_run(plan._h, psi.data_ptr(), lam.data_ptr(), mats.data_ptr(), grad.data_ptr(), arr,
                                           engine._stream(psi)))

    # local pieces of measure_dist (csrc/b200q_sample.cu)
    def block_mass(self, amps, nlocal):
        return engine.block_mass(amps, nlocal, 1)[0]

    def sample_indices(self, amps, nlocal, uniforms, mass):
        return engine.sample_indices(amps, nlocal, uniforms, mass=mass)

    def marginal_probs(self, amps, nlocal, mask, keys_sorted):
        return engine.marginal_probs(amps, nlocal, mask, keys_sorted)


class ShardedProgram:
    """Per-rank execution schedule of a lowered gate program: local segments separated by block transposes."""

    def __init__(self, low: Lowering, nqubit: int, world_size: int, rank: int, mode: str = 'pswap'):
        """mode 'pswap': exchanges are block transposes (rank bits <-> top local bits; the qubits to evict are first
        moved to the top local positions by relabelling SWAPs inside the fused passes) -- the NCCL path.
        mode 'perm': exchanges are arbitrary permutations of the bits of the distributed index, done by the fused
        last pass of the preceding segment (`b200q_plan_run_exchange`): victims leave from where they are, and
        the reference layout is restored by ONE permuting exchange instead of up to two transposes plus passes of
        relabelling SWAPs."""
        self.low, self.n, self.world, self.rank = low, nqubit, world_size, rank
        self.g = world_size.bit_length() - 1
        self.nl = nqubit - self.g
        self.mode = mode
        # 'perm' mode, B200Q_SHARD_TRIM=k: a segment is cut where its passes get sparse (fewer than k gates); the tail gates
        # run at the start of the next segment, fused with its abundant first gates, and the exchange rides on a DENSE
        # last pass.  Measured on 8 GPUs with k = 20: 30 qubits 38.7 ms (53 passes, 10 exchanges) against 40.0 ms (58, 9),
        # but 33 qubits 238.3 ms (45 passes) against 229.9 ms (48): an exchanging pass costs its NVLink time PLUS the
        # arithmetic it carries (the remote stores back-pressure the CTAs), so a dense one gives back what the saved
        # passes won.  Off by default.
        self.trim = int(os.environ.get('B200Q_SHARD_TRIM', '0')) if mode == 'perm' else 0
        self.n_deferred = 0
        self.plans = {}
        self._schedule()

    def _schedule(self):
        """Greedy list scheduling with commutation (same rule as the C++ planner) + Belady-style eviction."""
        n, g, nl = self.n, self.g, self.nl
        recs = self.low.records
        phys = list(range(n))                # logical bit -> physical bit
        done = [False] * len(recs)
        steps = []                           # ('seg', [localised records / ('pswap', a, b)]) | ('swap',)
        remaining = len(recs)
        nswaps = 0
        INF = 1 << 60

        def acts(rec):
            kind, targets, ctrl = rec[0], rec[1], rec[2]
            tmask = sum(1 << t for t in targets)
            dense_diag = kind == L.GATE_DIAG and len(targets) > 2   # lowered as a dense gate by the planner
            tm = tmask if (kind != L.GATE_DIAG or dense_diag) else 0
            dm = sum(1 << c for c in ctrl) | (tmask if tm == 0 else 0)
            return tm, dm

        def local_swaps(seg, wanted):
            """wanted: {physical position: logical bit}; emit physical SWAPs until satisfied."""
            for pos, b in wanted.items():
                if phys[b] != pos:
                    other = phys.index(pos)
                    seg.append(('pswap', phys[b], pos))
                    phys[other], phys[b] = phys[b], pos

        def transpose():
            nonlocal nswaps
            inv = {p: b for b, p in enumerate(phys)}
            for jj in range(g):
                a, b = inv[nl + jj], inv[nl - g + jj]
                phys[a], phys[b] = phys[b], phys[a]
            steps.append(('swap',))
            nswaps += 1

        def exchange_perm(new_phys):
            """perm mode: one exchange that moves every logical bit b from phys[b] to new_phys[b]."""
            nonlocal nswaps
            pm = list(range(n))
            for b in range(n):
                pm[phys[b]] = new_phys[b]
            assert pm[0] == 0, 'index bit 0 never moves (it lives inside a 16-byte complex64 chunk)'
            steps.append(('xperm', pm))
            phys[:] = new_phys
            nswaps += 1

        while remaining:
            glob_mask = sum(1 << b for b in range(n) if phys[b] >= nl)
            seg, bfull, bdiag, first_blocked = [], 0, 0, None
            for i, rec in enumerate(recs):
                if done[i]:
                    continue
                tm, dm = acts(rec)
                ok = not (tm & (bfull | bdiag)) and not (dm & bfull) and not (tm & glob_mask)
                if ok:
                    seg.append(self._localise(i, rec, phys))
                    done[i] = True
                    remaining -= 1
                else:
                    if first_blocked is None and (tm & glob_mask) and not (tm & (bfull | bdiag)) and not (dm & bfull):
                        first_blocked = tm
                    bfull |= tm
                    bdiag |= dm
            if remaining and self.trim and g > 0 and first_blocked is not None:
                seg, back = self._trim_segment(seg)
                for i in back:
                    done[i] = False
                remaining += len(back)
                self.n_deferred += len(back)
            if remaining:
                if g == 0 or first_blocked is None:
                    raise RuntimeError('sharded scheduler stalled (internal error)')
                # evict the g local qubits whose next non-diagonal use is farthest away (never a target of the
                # gate we are unblocking), move them to the top g local positions, then block-transpose
                next_use = [INF] * n
                for i in range(len(recs) - 1, -1, -1):
                    if not done[i]:
                        tm, _ = acts(recs[i])
                        for b in range(n):
                            if tm >> b & 1:
                                next_use[b] = i
                # perm mode: a victim leaves from where it is, so consecutive chunks of the source alternate between
                # destination ranks every 2^position amplitudes: only positions >= 8 (runs of >= 2 KiB) may leave;
                # the qubits below stay local for good (and index bit 0 lives inside a complex64 chunk)
                low = 0 if self.mode != 'perm' else (8 if nl >= 16 else 1)
                cand = [b for b in range(n) if low <= phys[b] < nl and not (first_blocked >> b & 1)]
                if len(cand) < g and low > 1:
                    cand = [b for b in range(n) if 1 <= phys[b] < nl and not (first_blocked >> b & 1)]
                if len(cand) < g:
                    raise RuntimeError('not enough local qubits to unblock a gate: it needs more than n_local - g targets')
                cand.sort(key=lambda b: (-next_use[b], -phys[b]))
                victims = cand[:g]
                if self.mode == 'perm':
                    # the g global qubits take the places of the g victims, wherever those are: no relabelling
                    if seg:
                        steps.append(('seg', seg))
                    inv = {p: b for b, p in enumerate(phys)}
                    new_phys = list(phys)
                    for jj, v in enumerate(victims):
                        gl = inv[nl + jj]
                        new_phys[gl], new_phys[v] = phys[v], phys[gl]
                    exchange_perm(new_phys)
                    continue
                # keep victims that already sit in the top positions where they are
                tops = [nl - g + jj for jj in range(g)]
                placed = {phys[b]: b for b in victims if phys[b] in tops}
                free = [t for t in tops if t not in placed]
                wanted = dict(placed)
                for b in victims:
                    if phys[b] not in tops:
                        wanted[free.pop()] = b
                local_swaps(seg, wanted)
            if seg:
                steps.append(('seg', seg))
            if remaining:
                transpose()
        # ---- restore the reference layout (logical bit b at physical bit b) -------------------------------
        if self.mode == 'perm' and phys != list(range(n)):
            exchange_perm(list(range(n)))
        if phys != list(range(n)):
            should_glob = list(range(nl, n))
            if any(phys[b] != b for b in should_glob):
                seg = []
                if any(phys[b] >= nl for b in should_glob):
                    # bring every currently-global qubit in, evicting only qubits that belong to the local part
                    cand = [b for b in range(nl) if phys[b] < nl]
                    cand.sort(key=lambda b: -phys[b])
                    tops = [nl - g + jj for jj in range(g)]
                    victims = cand[:g]
                    placed = {phys[b]: b for b in victims if phys[b] in tops}
                    free = [t for t in tops if t not in placed]
                    wanted = dict(placed)
                    for b in victims:
                        if phys[b] not in tops:
                            wanted[free.pop()] = b
                    local_swaps(seg, wanted)
                    if seg:
                        steps.append(('seg', seg))
                    transpose()
                    seg = []
                local_swaps(seg, {nl - g + jj: nl + jj for jj in range(g)})
                if seg:
                    steps.append(('seg', seg))
                transpose()
            seg = []
            local_swaps(seg, {b: b for b in range(nl)})
            if seg:
                steps.append(('seg', seg))
        assert phys == list(range(n)), phys
        self.steps, self.n_swaps = steps, nswaps
        self.n_segments = sum(1 for s in steps if s[0] == 'seg')

    def _trim_segment(self, seg):
        """Cut a segment that is followed by an exchange where its passes get sparse.  The gates that can run without a
        global qubit thin out towards the end of a segment (long dependency chains), and the fused planner then spends
        whole passes over the shard on a handful of gates -- the last of them, with 1-3 gates, carries the exchange.
        The segment is planned once with rank-independent stand-ins (every gate active, rank selectors dropped), kept up
        to its last pass with at least `self.trim` gates, and the gates of the later passes go back to the pending list:
        they run at the start of the next segment.  Returns (kept entries, record indices to put back).  The decision
        uses nothing rank-dependent, so every rank cuts at the same gate."""
        if len(seg) < 2 * self.trim:
            return seg, []
        proxies, owner = [], []
        for k, (i, kind, pt, ctrl, adj, active, fix, hint) in enumerate(seg):
            if fix is not None:
                kind, hint = L.GATE_DIAG, 0
            if len(pt) == 0:
                if not ctrl:
                    continue                 # a pure per-rank phase: no local bit, commutes with everything local
                pt, ctrl = (ctrl[0],), tuple(ctrl[1:])
            proxies.append(L.make_gate(kind, pt, ctrl, 0, adj, hint))
            owner.append(k)
        plan = engine.FusedPlan(self.nl, torch.complex64, proxies)
        counts = [plan.pass_gates(p) for p in range(plan.n_passes)]
        dense = [p for p, c in enumerate(counts) if c >= self.trim]
        if not dense or dense[-1] == len(counts) - 1:
            return seg, []
        drop = set()
        for p in range(dense[-1] + 1, len(counts)):
            drop.update(owner[q] for q in plan.pass_gate_ids(p))
        return [e for k, e in enumerate(seg) if k not in drop], [seg[k][0] for k in sorted(drop)]

    def _localise(self, i, rec, phys):
        """Rewrite record i for this rank: physical local bits, global controls resolved against the rank,
        global diagonal selectors folded into a per-rank derived matrix."""
        kind, targets, ctrl, adj, block, idx, size, hint = rec
        nl, rank = self.nl, self.rank
        local_ctrl, active = [], True
        for c in ctrl:
            p = phys[c]
            if p >= nl:
                if not (rank >> (p - nl)) & 1:
                    active = False
            else:
                local_ctrl.append(p)
        pt = [phys[t] for t in targets]
        fix = None
        if kind == L.GATE_DIAG and any(p >= nl for p in pt):
            # selector bits on rank bits are constants for this rank
            fix = [(j, (rank >> (p - nl)) & 1) for j, p in enumerate(pt) if p >= nl]
            pt = [p for p in pt if p < nl]
        return (i, kind, tuple(pt), tuple(local_ctrl), adj, active, fix, hint)

    def _segment_structs(self, step, mats, needs=None):
        """GateStruct list (+ extra derived matrices) of one local segment for this rank.  `needs`: optional list that
        receives, per struct, whether its matrix is computed from parameters / data (the reverse sweep accumulates
        cotangents only for those)."""
        structs, extra, off_extra = [], [], mats.numel()
        for item in step[1]:
            if item[0] == 'pswap':       # physical SWAP of two local bits = three CX relabellings
                a, b = item[1], item[2]
                structs += [L.make_gate(L.GATE_X, [b], [a]), L.make_gate(L.GATE_X, [a], [b]),
                            L.make_gate(L.GATE_X, [b], [a])]
                if needs is not None:
                    needs += [0, 0, 0]
                continue
            (i, kind, pt, ctrl, adj, active, fix, hint) = item
            if not active:
                continue
            if needs is not None:
                rec = self.low.records[i]
                block = self.low.derived[rec[5]][4] if rec[4] == 'derived' else rec[4]
                needs.append(0 if block in ('none', 'const') else 1)
            off = self.low.offsets[i]
            if fix is not None:
                k = len(self.low.records[i][1])
                d = torch.diagonal(mats[off:off + 4**k].reshape(2**k, 2**k))
                sel = d.reshape([2] * k)         # index order: matrix bit k-1 ... bit 0
                for j, bit in sorted(fix):   # ascending j = last axis first, so earlier axes keep their index
                    sel = sel.select(k - 1 - j, bit)
                vals = sel.reshape(-1)
                if len(pt) == 0:
                    # a pure per-rank phase v (all selectors are rank bits): without local controls a diagonal gate
                    # diag(v, v) on local bit 0; with local controls the first control becomes the target of
                    # diag(1, v) and the others stay controls (a placeholder target could collide with a control)
                    if ctrl:
                        pt, ctrl = (ctrl[0],), tuple(ctrl[1:])
                        vals = torch.cat([torch.ones_like(vals), vals])
                    else:
                        pt, vals = (0,), vals.repeat(2)
                extra.append(torch.diag(vals).reshape(-1))
                structs.append(L.make_gate(L.GATE_DIAG, pt, ctrl, off_extra, adj))
                off_extra += extra[-1].numel()
            else:
                structs.append(L.make_gate(kind, pt, ctrl, off, adj, hint))
        return structs, extra

    def run_step(self, si: int, state, mats: torch.Tensor, executor, fuse_exchange: bool) -> str:
        """Execute step `si` for this rank.  Returns 'exchange' when a fused pass + exchange was issued: the caller
        must order the ranks and then call `commit_exchange(state)` (shard and receive buffer swap roles)."""
        step = self.steps[si]
        nl = self.nl
        if step[0] in ('swap', 'xperm') and self._skip_next:     # done by the fused last pass of the previous segment
            self._skip_next = False
            return 'skipped'
        if step[0] == 'swap':
            block_transpose(state.amps, state.buffer)
            state.amps, state.buffer = state.buffer, state.amps
            return 'swap'
        if step[0] == 'xperm':
            # an exchange with no segment in front of it: the identity as a one-op plan, fused with the exchange
            assert fuse_exchange, "the 'perm' schedule needs the peer-mapped exchange"
            key = ('identity', state.amps.dtype)
            if key not in self.plans:
                self.plans[key] = executor.make_plan(nl, state.amps.dtype, [L.make_gate(L.GATE_DIAG, (1,), (), 0)],
                                                     exchange=True)
            eye = torch.eye(2, dtype=state.amps.dtype, device=state.amps.device).reshape(-1)
            executor.run_plan_exchange(self.plans[key], state.amps, eye, state.peer_buffer_ptrs(), self.rank, step[1])
            self.fused_exchanges += 1
            return 'exchange'
        structs, extra = self._segment_structs(step, mats)
        if not structs:
            return 'empty'
        m = torch.cat([mats] + extra) if extra else mats
        nxt = self.steps[si + 1] if si + 1 < len(self.steps) else None
        fused = fuse_exchange and nxt is not None and nxt[0] in ('swap', 'xperm')
        if fused and any((g.n_targets > 4 or (g.n_targets > 2 and g.flags & L.GATE_GRAD)) and g.kind != L.GATE_X
                         for g in structs):
            fused = False    # a dense gate on 5-6 targets is a pass of its own and cannot carry the exchange scatter:
            #                  the exchange then runs as a stand-alone (identity plan) exchange step
        key = (si, state.amps.dtype, fused)
        if key not in self.plans:
            self.plans[key] = (executor.make_plan(nl, state.amps.dtype, structs, exchange=True) if fused
                               else executor.make_plan(nl, state.amps.dtype, structs))
        if fused:
            # fused pass + exchange: the segment's last pass stores into the peers' receive buffers over NVLink
            executor.run_plan_exchange(self.plans[key], state.amps, m, state.peer_buffer_ptrs(), self.rank,
                                       nxt[1] if nxt[0] == 'xperm' else None)
            self._skip_next = True
            self.fused_exchanges += 1
            return 'exchange'
        executor.run_plan(self.plans[key], state.amps, m)
        return 'seg'

    @staticmethod
    def commit_exchange(state) -> None:
        state.amps, state.buffer = state.buffer, state.amps

    def run(self, state: DistributedQubitState, mats: torch.Tensor, executor, marks=None) -> None:
        """`marks`: optional list receiving (kind, start_event, end_event) per step (bench.py timing)."""
        import torch.distributed as dist
        # only the 'perm' schedule fuses: with block transposes a rank whose segment is empty (all its gates
        # switched off by global controls) would enter the NCCL transpose while its peers store directly
        fuse_exchange = (self.mode == 'perm' and hasattr(executor, 'run_plan_exchange') and self.world > 1
                         and state.enable_peer_exchange())
        assert fuse_exchange or self.mode != 'perm', "the 'perm' schedule needs the peer-mapped exchange"
        self.fused_exchanges = 0
        self._skip_next = False
        for si in range(len(self.steps)):
            ev0 = None
            if marks is not None:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record()
            what = self.run_step(si, state, mats, executor, fuse_exchange)
            if what == 'exchange':
                # a tiny all-reduce orders the ranks (all peer stores are complete when the kernels have finished
                # everywhere), then shard and receive buffer swap roles
                dist.all_reduce(self._sync_token(state.amps.device))
                self.commit_exchange(state)
            if marks is not None and what in ('swap', 'seg', 'exchange'):
                ev1 = torch.cuda.Event(enable_timing=True)
                ev1.record()
                marks.append(('swap' if what == 'swap' else 'seg', ev0, ev1))

    def run_adjoint(self, psi: DistributedQubitState, lam: DistributedQubitState, mats: torch.Tensor, executor) -> torch.Tensor:
        """Reverse sweep over the sharded schedule (reference adjoint.py:47-83 generalised to the full matrix
        cotangent): `psi` enters as the final state and leaves as the initial one, `lam` enters as the cotangent of the
        final state.  Exchanges are bit permutations of the distributed index: they are replayed backwards on both
        states; every local segment runs the reverse-sweep kernel on the two shards.  Returns this rank's share of
        the cotangent of `mats` (complex, like `mats`); the caller all-reduces it."""
        import torch.distributed as dist
        nl = self.nl
        fuse_exchange = (self.mode == 'perm' and hasattr(executor, 'run_plan_exchange') and self.world > 1)
        leaf = mats.detach().clone().requires_grad_(True)
        total = torch.zeros_like(leaf)
        for si in range(len(self.steps) - 1, -1, -1):
            step = self.steps[si]
            if step[0] == 'swap':            # block transpose: its own inverse
                for st in (psi, lam):
                    block_transpose(st.amps, st.buffer)
                    st.amps, st.buffer = st.buffer, st.amps
                continue
            if step[0] == 'xperm':
                assert fuse_exchange, "the 'perm' schedule needs the peer-mapped exchange"
                pm = step[1]
                inv = [0] * len(pm)
                for j, pj in enumerate(pm):
                    inv[pj] = j
                key = ('identity', psi.amps.dtype)
                if key not in self.plans:
                    self.plans[key] = executor.make_plan(nl, psi.amps.dtype, [L.make_gate(L.GATE_DIAG, (1,), (), 0)],
                                                         exchange=True)
                eye = torch.eye(2, dtype=psi.amps.dtype, device=psi.amps.device).reshape(-1)
                for st in (psi, lam):
                    executor.run_plan_exchange(self.plans[key], st.amps, eye, st.peer_buffer_ptrs(), self.rank, inv)
                    dist.all_reduce(self._sync_token(st.amps.device))
                    self.commit_exchange(st)
                continue
            needs = []
            with torch.enable_grad():
                structs, extra = self._segment_structs(step, leaf, needs)
                if not structs:
                    continue
                m = torch.cat([leaf] + extra) if extra else leaf
            key = (si, psi.amps.dtype, 'adjoint')
            if key not in self.plans:
                self.plans[key] = executor.make_plan(nl, psi.amps.dtype, structs)
            grad = torch.zeros(m.numel(), dtype=torch.complex128, device=m.device)
            executor.run_plan_adjoint(self.plans[key], psi.amps, lam.amps, m.detach().contiguous(), grad, needs)
            if extra:        # chain the derived per-rank matrices (global diagonal selectors) back to the buffer
                (g_leaf,) = torch.autograd.grad(m, leaf, grad.to(m.dtype))
                total += g_leaf
            else:
                total += grad.to(m.dtype)
        return total

    def _sync_token(self, device):
        tok = self.__dict__.get('_tok')
        if tok is None or tok.device != device:
            tok = torch.zeros(1, dtype=torch.int32, device=device)
            self.__dict__['_tok'] = tok
        return tok

    def stats(self):
        """(gates, local passes, segments, block transposes) of this rank's schedule (after a first run)."""
        passes = sum(getattr(p, 'n_passes', 0) for p in self.plans.values())
        return {'gates': len(self.low.records), 'passes': passes, 'segments': self.n_segments, 'swaps': self.n_swaps}


def measure_dist(state: DistributedQubitState, shots: int = 1024, with_prob: bool = False, wires=None,
                 block_size: int = 2**24, executor=None, generator: torch.Generator | None = None) -> dict:
    """Measure a sharded statevector (reference distributed.py:205-285).  Returns the result dict on rank 0 and
    `{}` on the other ranks, like the reference.

    No amplitudes move: every rank reduces its shard to block masses (one read), the W shard totals are
    all-gathered, rank 0 draws the uniforms and broadcasts them, every rank runs the inverse CDF for the shots
    that fall into its own mass interval, and only (key, count[, probability]) pairs are gathered.  The
    reference instead swaps measured global wires into the shard (`dist_swap_gate`) and all-reduces a 2^k
    probability tensor."""
    import torch.distributed as dist
    from .qmath import measure as _measure
    ex = executor or CudaExecutor()
    if state.world_size == 1 and executor is None:
        return _measure(state.amps, shots, with_prob, wires, False, block_size, generator=generator)
    n, nl, rank, world = state.nqubit, state.log_num_amps_per_node, state.rank, state.world_size
    if isinstance(wires, int):
        wires = [wires]
    meas = list(range(n)) if wires is None else sorted(wires)
    nbits = len(meas)
    amps = state.amps
    dev = amps.device
    mass = ex.block_mass(amps, nl)
    totals = torch.zeros(world, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_gather_into_tensor(totals, mass.sum().reshape(1))
    else:
        totals[0] = mass.sum()
    u = torch.rand(shots, dtype=torch.float64, generator=generator).to(dev)
    if world > 1:
        dist.broadcast(u, src=0)
    cdf = torch.cumsum(totals, 0)
    t = u * cdf[-1]
    owner = torch.searchsorted(cdf, t, right=True).clamp_(max=world - 1)
    mine = owner == rank
    local = {}
    if bool(mine.any()):
        before = cdf[rank] - totals[rank]
        ul = ((t[mine] - before) / totals[rank]).clamp_(0.0, 1.0 - 2.0**-53)
        idx = ex.sample_indices(amps, nl, ul, mass) | (rank << nl)
        keys = idx if nbits == n else torch.zeros_like(idx)
        if nbits != n:
            for j, w in enumerate(meas):
                keys |= ((idx >> (n - 1 - w)) & 1) << (nbits - 1 - j)
        vals, counts = torch.unique(keys, return_counts=True)
        local = dict(zip(vals.tolist(), counts.tolist()))
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, local)
    else:
        gathered = [local]
    merged = {}
    for d in gathered:
        for k, c in d.items():
            merged[k] = merged.get(k, 0) + c
    probs = None
    if with_prob:
        allkeys = sorted(merged)
        kt = torch.tensor(allkeys, dtype=torch.int64, device=dev)
        dep = torch.zeros_like(kt)        # key bits deposited at the measured GLOBAL index bits
        mask = 0
        for j, w in enumerate(meas):
            dep |= ((kt >> (nbits - 1 - j)) & 1) << (n - 1 - w)
            mask |= 1 << (n - 1 - w)
        lmask = mask & ((1 << nl) - 1)
        gmask = mask >> nl                                # measured rank bits
        on_rank = ((dep >> nl) & gmask) == (rank & gmask)
        part = torch.zeros(len(allkeys), dtype=torch.float64, device=dev)
        if bool(on_rank.any()):
            ldep = (dep[on_rank] & lmask)
            uniq, inv = torch.unique(ldep, return_inverse=True)      # sorted; several keys may share a local part
            part[on_rank] = ex.marginal_probs(amps, nl, lmask, uniq.contiguous())[inv]
        if world > 1:
            dist.all_reduce(part)
        probs = dict(zip(allkeys, part.tolist()))
    if rank != 0:
        return {}
    out = {}
    for k in sorted(merged):
        key = format(k, f'0{nbits}b')
        out[key] = (merged[k], probs[k]) if with_prob else merged[k]
    return out
