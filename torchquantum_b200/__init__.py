"""torchquantum_b200 — a B200-native (sm_100a) statevector engine behind TorchQuantum's
gate-application API: `tq.QuantumDevice`, `tqf.<gate>` / `qdev.<gate>`, `tq.Operator`,
`tq.QuantumModule`, `expval_joint_analytical`, `expval` / `MeasureAll`, `measure`.

    import torchquantum_b200 as tq
    import torchquantum_b200.functional as tqf
"""
from .macro import *  # noqa: F401,F403
from .module import QuantumModule, QuantumModuleList, QuantumModuleDict, QuantumGraph, static_support  # noqa: F401
from .device import QuantumDevice  # noqa: F401
from .sharded import ShardedQuantumDevice  # noqa: F401
from . import functional  # noqa: F401
from . import operator  # noqa: F401
from .operator import *  # noqa: F401,F403
from .operator import op_name_dict  # noqa: F401
from .measurement import *  # noqa: F401,F403
from .measure_modules import MeasureAll, MeasureMultipleTimes, MeasureMultiPauliSum, MeasureMultiQubitPauliSum  # noqa: F401
from .encoding import *  # noqa: F401,F403
from .layer import *  # noqa: F401,F403
from .graphs import GraphedStep  # noqa: F401
from .algorithm import Hamiltonian, VQE  # noqa: F401
from .op_hamil_exp import OpHamilExp, OpPauliExp  # noqa: F401
from .gradients import parameter_shift_gradient, hadamard_grad  # noqa: F401

operator.OpHamilExp, operator.OpPauliExp = OpHamilExp, OpPauliExp   # reference location: torchquantum.operator

__version__ = "0.1.0"
