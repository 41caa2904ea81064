"""OpenQASM 3.0 <-> QubitCircuit (reference src/deepquantum/qasm3.py:40-156 export, :166-472 import), so that
external benchmark circuits reach the fused kernels.

Same dialect as the reference: one register `qubit[n] q;`, the stdgates names `u p x y z h s sdg t tdg rx ry rz
swap cx cz ccx cswap rxx ryy rzz`, `barrier`, `c[i] = measure q[i];`, user gates (`def name(params) a, b { ... }`, also
accepted with the standard keyword `gate`), and the modifiers `inv @`, `ctrl @` (repeatable) and `pow(k) @` with
integer or fractional, positive or negative `k`.  Written as a statement parser (comments stripped, statements split
on `;` and braces) with a small arithmetic evaluator instead of `eval`; host-only -- nothing here touches the device.
"""
from __future__ import annotations

import ast
import math
import operator
import re

import torch

from .circuit import QubitCircuit
from .gate import (Barrier, CNOT, Fredkin, Hadamard, PauliX, PauliY, PauliZ, PhaseShift, Rx, Rxx, Ry, Ryy, Rz, Rzz,
                   SDaggerGate, SGate, Swap, TDaggerGate, TGate, Toffoli, U3Gate)
from .operation import Channel, Gate, Layer

# ---------------------------------------------------------------------------------------------------------------
# export
# ---------------------------------------------------------------------------------------------------------------
_EXPORT_NAMES = {U3Gate: 'u', PhaseShift: 'p', PauliX: 'x', PauliY: 'y', PauliZ: 'z', Hadamard: 'h', SGate: 's',
                 SDaggerGate: 'sdg', TGate: 't', TDaggerGate: 'tdg', Rx: 'rx', Ry: 'ry', Rz: 'rz', Swap: 'swap',
                 CNOT: 'cx', Toffoli: 'ccx', Fredkin: 'cswap', Rxx: 'rxx', Ryy: 'ryy', Rzz: 'rzz'}


def _statement_of(op) -> str:
    if isinstance(op, Layer):
        return '\n'.join(_statement_of(g) for g in op.gates)
    if isinstance(op, Barrier):
        return 'barrier ' + ', '.join(f'q[{w}]' for w in op.wires) + ';'
    if isinstance(op, Channel):
        return f'// Quantum channels like {op.name} are not part of the OpenQASM 3.0 core specification.'
    if not isinstance(op, Gate):
        return f'// Unsupported operation type: {op.__class__.__name__}'
    name = _EXPORT_NAMES.get(type(op))
    if name is None:
        return f'// Unsupported gate: {op.name}'
    args = ''
    if op.npara > 0:
        values = [getattr(op, p).item() for p in op._pnames]
        if getattr(op, 'inv_mode', False):
            values = [-v for v in values]
        args = '(' + ', '.join(str(v) for v in values) + ')'
    if isinstance(op, (CNOT, Toffoli, Fredkin)):          # controls are part of the wires
        return f'{name} ' + ', '.join(f'q[{w}]' for w in op.wires) + ';'
    qubits = ', '.join(f'q[{w}]' for w in op.controls + op.wires)
    return 'ctrl @ ' * len(op.controls) + f'{name}{args} {qubits};'


def cir_to_qasm3(circuit: QubitCircuit) -> str:
    """OpenQASM 3.0 text of a circuit (reference qasm3.py:117-156)."""
    out = ['OPENQASM 3.0;', 'include "stdgates.inc";', f'qubit[{circuit.nqubit}] q;']
    if circuit.wires_measure:
        out.append(f'bit[{max(circuit.wires_measure) + 1}] c;')
    out += [line for line in (_statement_of(op) for op in circuit.operators) if line]
    if circuit.wires_measure:
        out.append('\n// Measurements')
        out += [f'c[{w}] = measure q[{w}];' for w in sorted(circuit.wires_measure)]
    return '\n'.join(out)


# ---------------------------------------------------------------------------------------------------------------
# import
# ---------------------------------------------------------------------------------------------------------------
_BINOPS = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
           ast.Pow: operator.pow, ast.Mod: operator.mod, ast.FloorDiv: operator.floordiv}
_FUNCS = {'sin': math.sin, 'cos': math.cos, 'tan': math.tan, 'exp': math.exp, 'ln': math.log, 'log': math.log,
          'sqrt': math.sqrt, 'arcsin': math.asin, 'arccos': math.acos, 'arctan': math.atan, 'abs': abs}
_CONSTS = {'pi': math.pi, 'π': math.pi, 'tau': 2 * math.pi, 'euler': math.e}


def _evaluate(text: str, scope: dict) -> float:
    """Arithmetic over numbers, `pi`, the enclosing gate's parameters and a few functions."""
    def walk(node):
        if isinstance(node, ast.Expression):
            return walk(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            return node.value
        if isinstance(node, ast.Name):
            if node.id in scope:
                return scope[node.id]
            if node.id in _CONSTS:
                return _CONSTS[node.id]
            raise ValueError(f'unknown identifier {node.id!r} in expression {text!r}')
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            v = walk(node.operand)
            return -v if isinstance(node.op, ast.USub) else v
        if isinstance(node, ast.BinOp) and type(node.op) in _BINOPS:
            return _BINOPS[type(node.op)](walk(node.left), walk(node.right))
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in _FUNCS:
            return _FUNCS[node.func.id](*[walk(a) for a in node.args])
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id == 'np' \
                and node.attr == 'pi':
            return math.pi
        raise ValueError(f'unsupported expression {text!r}')
    return float(walk(ast.parse(text.strip().replace('^', '**'), mode='eval')))


class _GateDef:
    def __init__(self, params, qubits, body):
        self.params, self.qubits, self.body = params, qubits, body


def _split_top(text: str, sep: str = ',') -> list[str]:
    """Split on `sep` outside parentheses / brackets."""
    parts, depth, cur = [], 0, ''
    for ch in text:
        if ch in '([':
            depth += 1
        elif ch in ')]':
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur.strip())
            cur = ''
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


def _statements(text: str):
    """Comment-free source -> list of ('def', header, [body statements]) / ('stmt', text)."""
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    text = '\n'.join(line.split('//')[0] for line in text.splitlines())
    out, i, n = [], 0, len(text)
    while i < n:
        m = re.compile(r'\s*(def|gate)\s+([^{;]*)\{').match(text, i)
        if m:
            depth, j = 1, m.end()
            while j < n and depth:
                depth += (text[j] == '{') - (text[j] == '}')
                j += 1
            body = [s.strip() for s in text[m.end():j - 1].split(';') if s.strip()]
            out.append(('def', m.group(2).strip(), body))
            i = j
            continue
        j = text.find(';', i)
        if j < 0:
            break
        stmt = text[i:j].strip()
        if stmt:
            out.append(('stmt', stmt))
        i = j + 1
    return out


_CALL = re.compile(r'^((?:(?:inv|ctrl|negctrl|pow\s*\([^@]*\))\s*@\s*)*)([A-Za-z_]\w*)'
                   r'(?:\s*\((.*)\)\s*|\s+)(.+)$', re.S)
_INVERSE_NAME = {'s': 'sdg', 'sdg': 's', 't': 'tdg', 'tdg': 't'}
_NEGATE = ('rx', 'ry', 'rz', 'p', 'rxx', 'ryy', 'rzz')
_PLAIN = {'h': 'h', 'x': 'x', 'y': 'y', 'z': 'z', 's': 's', 'sdg': 'sdg', 't': 't', 'tdg': 'tdg', 'swap': 'swap'}
_PARAMETRIC = {'rx': 'rx', 'ry': 'ry', 'rz': 'rz', 'p': 'p', 'u': 'u3', 'rxx': 'rxx', 'ryy': 'ryy', 'rzz': 'rzz'}
_BUILTINS = set(_PLAIN) | set(_PARAMETRIC) | {'cx', 'cz', 'ccx', 'cswap'}


def _local_unitary(circuit: QubitCircuit) -> torch.Tensor:
    """Dense matrix of a tiny circuit from its gate matrices (host-side matrix assembly, used only for fractional
    powers of a gate: the result becomes ONE `any` gate of the outer circuit)."""
    n = circuit.nqubit
    dim = 2**n
    total = torch.eye(dim, dtype=torch.cdouble)
    for op in circuit.operators:
        if isinstance(op, Barrier):
            continue
        local = op.update_matrix().detach().to(torch.cdouble).cpu()
        if type(op) in (CNOT, Toffoli, Fredkin) or not op.controls:
            wires, block = list(op.wires), local
        else:                                   # controlled gate: identity unless every control is 1
            wires = list(op.controls) + list(op.wires)
            block = torch.eye(2**len(wires), dtype=torch.cdouble)
            block[-local.shape[0]:, -local.shape[0]:] = local
        rest = [w for w in range(n) if w not in wires]
        perm = wires + rest
        full = torch.kron(block, torch.eye(2**len(rest), dtype=torch.cdouble)).reshape([2] * (2 * n))
        inv = [perm.index(w) for w in range(n)]
        full = full.permute(inv + [n + k for k in inv]).reshape(dim, dim)
        total = full @ total
    return total


def qasm3_to_cir(qasm_string: str) -> QubitCircuit:
    """Build a `QubitCircuit` from OpenQASM 3.0 text (reference qasm3.py:166-472)."""
    if 'OPENQASM 3' not in qasm_string:
        raise ValueError('Input is not a valid OpenQASM 3.0 string (Header missing).')
    items = _statements(qasm_string)
    definitions: dict[str, _GateDef] = {}
    main = []
    for item in items:
        if item[0] == 'def':
            m = re.match(r'^([A-Za-z_]\w*)\s*(?:\((.*?)\))?\s*(.*)$', item[1], re.S)
            name, params, qubits = m.group(1), m.group(2) or '', m.group(3) or ''
            definitions[name] = _GateDef([p.strip() for p in params.split(',') if p.strip()],
                                         [q.strip() for q in qubits.split(',') if q.strip()], item[2])
        else:
            main.append(item[1])
    nqubit = 0
    for stmt in main:
        m = re.match(r'^qubit\s*\[\s*(\d+)\s*\]', stmt)
        if m:
            nqubit = int(m.group(1))
            break
    if nqubit == 0:
        raise ValueError('Qubit declaration not found or zero qubits specified.')
    circuit = QubitCircuit(nqubit=nqubit)

    def index_of(q: str) -> int:
        m = re.match(r'^[A-Za-z_]\w*\s*\[\s*(\d+)\s*\]$', q.strip())
        if not m:
            raise ValueError(f'cannot resolve qubit operand {q!r}')
        return int(m.group(1))

    def builtin(cir, name, params, qubits, controls, inverted):
        if inverted:
            if name in _NEGATE:
                params = [-p for p in params]
            elif name == 'u':
                params = [-params[0], -params[2], -params[1]]
            name = _INVERSE_NAME.get(name, name)
        ctrl = list(controls)
        if name == 'cx':
            ctrl, target = ctrl + [qubits[0]], qubits[1]
            cir.cnot(ctrl[0], target) if len(ctrl) == 1 else cir.x(target, controls=ctrl)
        elif name == 'cz':
            cir.z(qubits[1], controls=ctrl + [qubits[0]])
        elif name == 'ccx':
            ctrl, target = ctrl + qubits[:2], qubits[2]
            cir.toffoli(ctrl[0], ctrl[1], target) if len(ctrl) == 2 else cir.x(target, controls=ctrl)
        elif name == 'cswap':
            ctrl, targets = ctrl + [qubits[0]], qubits[1:3]
            cir.fredkin(ctrl[0], targets[0], targets[1]) if len(ctrl) == 1 else cir.swap(targets, controls=ctrl)
        else:
            wires = qubits[0] if len(qubits) == 1 else qubits
            if name in _PLAIN:
                getattr(cir, _PLAIN[name])(wires, controls=ctrl)
            else:
                getattr(cir, _PARAMETRIC[name])(wires, params, controls=ctrl)

    def run(statements, cir, scope, controls, inverted, qubit_map):
        for stmt in (reversed(statements) if inverted else statements):
            head = re.match(r'^[A-Za-z_]\w*', stmt)
            head = head.group(0) if head else ''
            if head in ('OPENQASM', 'include', 'bit', 'qubit', 'defcal', 'const', 'input', 'output'):
                continue
            if re.search(r'\bmeasure\b', stmt):
                found = [int(w) for w in re.findall(r'\bmeasure\s+[A-Za-z_]\w*\s*\[\s*(\d+)\s*\]', stmt)]
                if not found and re.search(r'\bmeasure\s+[A-Za-z_]\w*\s*$', stmt):    # the whole register
                    found = list(range(cir.nqubit))
                for w in found:
                    if w not in cir.wires_measure:
                        cir.wires_measure.append(w)
                continue
            if head == 'barrier':
                ops = [o for o in _split_top(stmt[len('barrier'):]) if o]
                cir.barrier(wires=[index_of(qubit_map.get(o, o)) for o in ops] or None)
                continue
            m = _CALL.match(stmt)
            if not m:
                raise ValueError(f'cannot parse OpenQASM statement {stmt!r}')
            mods, name, arg_text, operand_text = m.groups()
            operands = [qubit_map.get(o, o) for o in _split_top(operand_text)]
            n_ctrl = len(re.findall(r'\bctrl\b', mods))
            if re.search(r'\bnegctrl\b', mods):
                raise ValueError('negctrl @ is not supported')
            flip = inverted ^ (len(re.findall(r'\binv\b', mods)) % 2 == 1)
            power = 1.0
            pm = re.search(r'pow\s*\((.*?)\)\s*@', mods, re.S)
            if pm:
                power = _evaluate(pm.group(1), scope)
            if flip:
                power = -power
            ctrl = controls + [index_of(o) for o in operands[:n_ctrl]]
            targets = operands[n_ctrl:]
            params = [_evaluate(a, scope) for a in _split_top(arg_text)] if arg_text and arg_text.strip() else []
            if name not in definitions and name not in _BUILTINS:
                raise ValueError(f'unsupported gate {name!r}')
            if power != int(power):                        # fractional power: eigen-decomposition of the block
                sub = QubitCircuit(len(targets))
                emit(name, params, list(range(len(targets))), sub, [], False)
                vals, vecs = torch.linalg.eig(_local_unitary(sub))
                block = vecs @ torch.diag(vals**power) @ torch.linalg.inv(vecs)
                cir.any(block.to(torch.cfloat), wires=[index_of(t) for t in targets], controls=ctrl)
                continue
            for _ in range(abs(int(power))):
                emit(name, params, [index_of(t) for t in targets], cir, ctrl, power < 0)

    def emit(name, params, targets, cir, ctrl, inverted):
        if name in definitions:
            d = definitions[name]
            if len(targets) != len(d.qubits) or len(params) != len(d.params):
                raise ValueError(f'wrong number of operands or parameters for gate {name!r}')
            qmap = {formal: f'q[{actual}]' for formal, actual in zip(d.qubits, targets)}
            run(d.body, cir, dict(zip(d.params, params)), ctrl, inverted, qmap)
        else:
            builtin(cir, name, params, targets, ctrl, inverted)

    run(main, circuit, {}, [], False, {})
    circuit.wires_measure.sort()
    return circuit


__all__ = ['cir_to_qasm3', 'qasm3_to_cir']
