"""torchquantum_b200 — a B200-native (sm_100a) statevector engine behind TorchQuantum's
gate-application API: `tq.QuantumDevice`, `tqf.<gate>` / `qdev.<gate>`, `tq.Operator`,
`tq.QuantumModule`, `expval_joint_analytical`, `expval` / `MeasureAll`, `measure`.

    import torchquantum_b200 as tq
    import torchquantum_b200.functional as tqf
"""
from .macro import *  # noqa: F401,F403
from .device import QuantumDevice  # noqa: F401
from . import functional  # noqa: F401

__version__ = "0.1.0"
