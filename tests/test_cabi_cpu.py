"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/rlmpc_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rlmpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rlmpc_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported(built_library):
    from mpc4rl_b200 import _cabi

    lib = C.CDLL(built_library)
    decl = _declared_symbols()
    assert len(decl) >= 15
    for s in decl:
        assert hasattr(lib, s), f"{s} declared in include/rlmpc_b200.h but not exported"
    assert sorted(_cabi.SYMBOLS) == decl


def test_desc_struct_matches_header(built_library):
    from mpc4rl_b200 import _cabi

    # struct rlmpc_problem_desc: 2 ints + (129 + 6*8 + 24 + 4*8) doubles
    assert C.sizeof(_cabi.ProblemDesc) == 8 + 8 * (129 + 48 + 24 + 32)


def test_no_gpu_means_loud_failure(built_library):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mpc4rl_b200 import BatchedMPC, cartpole_original_config, cartpole_spec

    with pytest.raises(RuntimeError):
        BatchedMPC(cartpole_spec(cartpole_original_config()), max_batch=4, device=0)
    from mpc4rl_b200 import _cabi

    lib = _cabi.load()
    h = C.c_void_p()
    d = cartpole_spec(cartpole_original_config()).to_desc()
    assert lib.rlmpc_create(C.byref(d), 4, 0, C.byref(h)) == -4  # RLMPC_ENODEV
    assert b"no CPU fallback" in lib.rlmpc_last_error()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under mpc4rl_b200/ may reference it."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "mpc4rl_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "oracle/" not in txt.replace("oracle/cpu_port (test infrastructure)", ""), f


def test_cartpole_spec_follows_reference_config():
    import numpy as np

    from mpc4rl_b200 import cartpole_original_config, cartpole_spec

    s = cartpole_spec(cartpole_original_config())
    assert (s.N, s.nx, s.nu, s.ntheta, s.np_model) == (40, 4, 1, 83, 3)
    assert np.isclose(s.model_const[0], 0.8 / 40 / 4)  # quirk Q1: one RK4 step of dT/4
    assert np.allclose(s.cost_scaling()[:40], 0.02) and s.cost_scaling()[40] == 1.0
    sl = s.p_slices()
    assert sl["W_0"][0] == slice(3, 28) and sl["yref_e"][0] == slice(79, 83)
    assert s.p_nominal[3] == 200.0 and s.p_nominal[3 + 6] == 0.02  # column-major diag


@pytest.mark.skipif(not os.path.isdir("/root/reference/config"), reason="reference checkout not present")
def test_builtin_configs_equal_reference_yaml():
    """cartpole_config()/cartpole_original_config() restate config/cartpole{,_original}.yaml; where the
    reference checkout is available, read the YAML files themselves and compare the resulting specs."""
    import numpy as np

    from mpc4rl_b200 import cartpole_config, cartpole_original_config, cartpole_spec
    from mpc4rl_b200.common.utils import read_config

    for name, mine in (("cartpole.yaml", cartpole_config()), ("cartpole_original.yaml", cartpole_original_config())):
        ref = read_config(os.path.join("/root/reference/config", name))["mpc"]
        a, b = cartpole_spec(ref), cartpole_spec(mine)
        assert (a.N, a.nx, a.nu, a.tf) == (b.N, b.nx, b.nu, b.tf), name
        for k in ("p_nominal", "lbu", "ubu", "lbx", "ubx", "lbx_e", "ubx_e", "model_const"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), (name, k)


def test_integration_md_struct_matches_the_header():
    """INTEGRATION.md shows the ctypes struct a maintainer would copy: it must have the size and field offsets of the
    binding the package itself uses (a stale copy under-allocates the struct and rlmpc_create reads past it)."""
    import ctypes as C  # noqa: F401  (used by the exec'd snippet)
    import re

    from mpc4rl_b200 import _cabi

    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"class ProblemDesc\(C\.Structure\):.*?\n\n", md, re.S)
    assert m, "INTEGRATION.md no longer shows the ProblemDesc struct"
    ns = {"C": C}
    exec(m.group(0), ns)
    doc = ns["ProblemDesc"]
    assert C.sizeof(doc) == C.sizeof(_cabi.ProblemDesc)
    assert [(n, getattr(doc, n).offset) for n, _ in doc._fields_] == [(n, getattr(_cabi.ProblemDesc, n).offset) for n, _ in _cabi.ProblemDesc._fields_]
    # and the header declares the same array lengths
    hdr = open(os.path.join(ROOT, "include", "rlmpc_b200.h")).read()
    assert "double model_const[24];" in hdr and "double lg[RLMPC_MAXD], ug[RLMPC_MAXD];" in hdr
