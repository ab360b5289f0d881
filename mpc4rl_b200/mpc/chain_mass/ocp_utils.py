"""Host-side helpers with the names of rlmpc/mpc/chain_mass/ocp_utils.py that the reference's drivers import
(examples/chain_mass.py:6-11): the parameter dictionary, the parameter-struct labels and the label search."""
from __future__ import annotations

from typing import List, Tuple

from ...problems import chain_define_x0, chain_mass_spec, chain_param_layout, get_chain_params  # noqa: F401


def define_nx_nu(n_mass: int) -> Tuple[int, int]:
    """ocp_utils.py:344-350"""
    return (2 * (n_mass - 2) + 1) * 3, 3


class _ParamStruct:
    """What the drivers use of define_param_struct_symSX(n_mass).cat: its entry labels (ocp_utils.py:353-371)."""

    def __init__(self, n_mass: int):
        _, self.size, self.labels = chain_param_layout(n_mass)
        self.cat = self

    def str(self) -> str:
        return "[" + ", ".join(self.labels) + "]"


def define_param_struct_symSX(n_mass: int, disturbance: bool = True) -> _ParamStruct:
    if not disturbance:
        raise NotImplementedError("the OCP of the reference always carries the disturbance parameters (ocp_utils.py:214)")
    return _ParamStruct(n_mass)


def find_idx_for_labels(sub_vars, sub_label: str) -> List[int]:
    """Indices whose label contains sub_label (ocp_utils.py:374-376)."""
    return [i for i, label in enumerate(sub_vars.str().strip("[]").split(", ")) if sub_label in label]
