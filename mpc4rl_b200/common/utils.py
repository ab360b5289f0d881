"""Small helpers mirroring rlmpc/common/utils.py of the reference."""
import os

import yaml

# acados multiplier order within one stage (rlmpc/common/utils.py:4-25)
ACADOS_MULTIPLIER_ORDER = [
    "lbu", "lbx", "lg", "lh", "lphi", "ubu", "ubx", "ug", "uh", "uphi",
    "lsbu", "lsbx", "lsg", "lsh", "lsphi", "usbu", "usbx", "usg", "ush", "usphi",
]


def read_config(config_file: str) -> dict:
    """YAML -> nested dict (rlmpc/common/utils.py:41-47)."""
    with open(config_file, "r") as stream:
        return yaml.safe_load(stream)


def get_root_path() -> str:
    return os.path.dirname(os.path.dirname(os.path.dirname(os.path.realpath(__file__))))
