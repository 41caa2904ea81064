"""deepquantum_b200 -- the DeepQuantum statevector gate-application path as hand-written sm_100a kernels.

Same names as the reference package for everything on that path (`QubitCircuit`, the gate and layer
classes, `QubitState`, `evolve_state`, `setup_distributed`, `DistributedQubitCircuit`), so
``import deepquantum_b200 as dq`` is a drop-in for circuits that stay on the statevector path.
Importing the package does not need the GPU; the first compute call loads `lib/libb200q.so` and
fails loudly if it (or an sm_100 device) is missing -- there is no CPU fallback.
"""
__version__ = '0.1.0'

from . import _lib, channel, engine, workloads  # noqa: F401
from ._lib import B200QError  # noqa: F401
from .circuit import DistributedQubitCircuit, QubitCircuit  # noqa: F401
from .communication import (cleanup_distributed, comm_exchange_arrays, comm_get_rank, comm_get_world_size,  # noqa: F401
                            setup_distributed)
from .distributed import DistributedQubitState  # noqa: F401
from .gate import (Barrier, CNOT, CombinedSingleGate, Fredkin, Hadamard, HamiltonianGate, Identity, ImaginarySwap,  # noqa: F401
                   LatentGate,
                   PauliX, PauliY, PauliZ, PhaseShift, ProjectionJ, ReconfigurableBeamSplitter, Rx, Rxx, Rxy, Ry, Ryy, Rz, Rzz,
                   Reset, SDaggerGate, SGate, Swap, TDaggerGate, TGate, Toffoli, U3Gate, UAnyGate)
from .layer import (CnotLayer, CnotRing, HLayer, Observable, RxLayer, RyLayer, RzLayer, U3Layer, XLayer,  # noqa: F401
                    YLayer, ZLayer)
from .channel import (AmplitudeDamping, BitFlip, Depolarizing, GeneralizedAmplitudeDamping, PhaseDamping,  # noqa: F401
                      PhaseFlip)
from .operation import Channel, Gate, Layer, Operation, dtype_map  # noqa: F401
from .qmath import (evolve_state, evolve_state_controlled, expectation, inverse_permutation, measure,  # noqa: F401
                    multi_kron, sample2expval)
from .state import QubitState, amplitude_encoding  # noqa: F401
from . import qasm3  # noqa: F401,E402
from .qasm3 import cir_to_qasm3, qasm3_to_cir  # noqa: F401,E402
from . import photonic  # noqa: F401,E402
from .photonic import QumodeCircuit  # noqa: F401,E402
from .photonic_distributed import DistributedFockState, DistributedQumodeCircuit  # noqa: F401,E402
