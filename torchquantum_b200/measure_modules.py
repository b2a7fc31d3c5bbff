"""Measurement modules — drop-in for measurement/measurements.py:315-456."""
import torch

from .device import QuantumDevice
from .measurement import expval
from .module import QuantumModule
from . import operator as op

__all__ = ["MeasureAll", "MeasureMultipleTimes", "MeasureMultiPauliSum", "MeasureMultiQubitPauliSum"]


def _permute(x, v_c_reg_mapping):
    if v_c_reg_mapping is None:
        return x
    c2v = v_c_reg_mapping["c2v"]
    perm = [c2v[k] for k in range(x.shape[-1]) if k in c2v.keys()]
    return x[:, perm]


class MeasureAll(QuantumModule):
    """Expectation value of `obs` on every wire -> [bsz, n_wires] (measurements.py:315-343)."""

    def __init__(self, obs, v_c_reg_mapping=None):
        super().__init__()
        self.obs = obs
        self.v_c_reg_mapping = v_c_reg_mapping

    def forward(self, qdev):
        # the reference builds `self.obs()` per call (measurements.py:331-336); a parameter-free observable is only asked for
        # its diagonalizing gates and eigenvalues, so one instance per module serves every step
        cached = self.__dict__.get("_obs_instance")
        if cached is not None and cached[0] is self.obs:
            observable = cached[1]
        else:
            observable = self.obs()
            if getattr(observable, "params", None) is None:
                self.__dict__["_obs_instance"] = (self.obs, observable)
        x = expval(qdev, list(range(qdev.n_wires)), [observable] * qdev.n_wires)
        x = _permute(x, self.v_c_reg_mapping)
        if self.noise_model_tq is not None and self.noise_model_tq.is_add_noise:
            return self.noise_model_tq.apply_readout_error(x)
        return x

    def set_v_c_reg_mapping(self, mapping):
        self.v_c_reg_mapping = mapping


class MeasureMultipleTimes(QuantumModule):
    """obs_list: [{'wires': [0, 2, 3, 1], 'observables': ['x', 'y', 'z', 'i']}, ...]; each entry is
    measured on a fresh clone of the device; results are concatenated along dim 0."""

    def __init__(self, obs_list, v_c_reg_mapping=None):
        super().__init__()
        self.obs_list = obs_list
        self.v_c_reg_mapping = v_c_reg_mapping

    def forward(self, qdev):
        res_all = []
        for layer in self.obs_list:
            qdev_new = QuantumDevice(n_wires=qdev.n_wires, device=qdev.state.device)
            qdev_new.clone_states(existing_states=qdev.states)
            observables = [op.I() for _ in range(qdev.n_wires)]
            for wire, observable in zip(layer["wires"], layer["observables"]):
                observables[wire] = op.op_name_dict[observable]()
            res = expval(qdev_new, wires=list(range(qdev.n_wires)), observables=observables)
            res_all.append(_permute(res, self.v_c_reg_mapping))
        return torch.cat(res_all)

    def set_v_c_reg_mapping(self, mapping):
        self.v_c_reg_mapping = mapping


class MeasureMultiPauliSum(QuantumModule):
    def __init__(self, obs_list, v_c_reg_mapping=None):
        super().__init__()
        self.obs_list = obs_list
        self.v_c_reg_mapping = v_c_reg_mapping
        self.measure_multiple_times = MeasureMultipleTimes(obs_list=obs_list, v_c_reg_mapping=v_c_reg_mapping)

    def forward(self, qdev):
        return self.measure_multiple_times(qdev).sum(-1)


class MeasureMultiQubitPauliSum(QuantumModule):
    """obs_list[0] = {'coefficient': [...]}, then one entry per term."""

    def __init__(self, obs_list, v_c_reg_mapping=None):
        super().__init__()
        self.obs_list = obs_list
        self.v_c_reg_mapping = v_c_reg_mapping
        self.measure_multiple_times = MeasureMultipleTimes(obs_list=obs_list[1:], v_c_reg_mapping=v_c_reg_mapping)

    def forward(self, qdev):
        res_all = self.measure_multiple_times(qdev)
        return (res_all * self.obs_list[0]["coefficient"]).sum(-1)
