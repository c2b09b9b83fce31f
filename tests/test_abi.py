"""The C-ABI boundary: the library loads, exports every symbol include/b200optas.h declares, the
Python structures match the header, and -- on a GPU-less machine -- compute calls fail loudly."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "b200optas.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bo_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from optas_b200 import _capi

    lib = _capi.load()
    names = _declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in b200optas.h but not exported"
    assert sorted(_capi.EXPORTS) == names
    assert lib.bo_abi_version() == _capi.BO_ABI_VERSION


def test_opcode_header_matches_front_end():
    import optas_b200.sym as S

    text = open(os.path.join(ROOT, "include", "bo_opcodes.h")).read()
    header = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define BO_(OP_\w+) (\d+)", text)}
    front = {k: v for k, v in vars(S).items() if k.startswith("OP_") and isinstance(v, int)}
    assert header == front


def test_options_struct_layout():
    from optas_b200 import _capi

    # flags,max_iter (8) + 4 doubles (32) + 2 pointers (16) + tpb, max_trips, 6 reserved (32)
    assert C.sizeof(_capi.bo_options) == 88
    assert C.sizeof(_capi.bo_tape) == 64
    assert C.sizeof(_capi.bo_sparsity) == 24


def test_compile_only_generates_sm100a_kernels():
    """Codegen + NVRTC for sm_100a works without a GPU and reports the kernel's resources."""
    import optas_b200
    from optas_b200 import problems
    from optas_b200.function import B200Function

    prob = problems.lwr_ik()
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True)
    info = solver.kernel_info()
    assert 0 < info["registers"] <= 255
    src = solver.kernel_source()
    # small dense problems run on the team tier: tape slices per warp, state in shared memory, no thread-local state
    assert solver.tier_info()["tier"] == "team"
    assert "bo_kkt_r0" in src and "bo_kkt_r3" in src and '#include "bo_ipm_team.cuh"' in src
    assert info["local_bytes"] <= 64
    old = optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True, team=False)
    assert old.tier_info()["tier"] == "dense" and '#include "bo_ipm_reg.cuh"' in old.kernel_source()
    fk = B200Function(prob.functions["fk_jac"], compile_only=True)
    assert 0 < fk.kernel_info()["registers"] <= 255
    assert "bo_sincos" in fk.kernel_source()


def test_malformed_tape_is_rejected():
    from optas_b200 import _capi
    from optas_b200.tape import Tape

    bad = Tape(instr=np.array([[12, 0, 5, 6]], dtype=np.int32), consts=np.zeros(0), n_work=1, in_sizes=[1], out_sizes=[1])
    with pytest.raises(_capi.BoError) as e:
        _capi.FunctionHandle(bad, flags=_capi.BO_FLAG_COMPILE_ONLY)
    assert e.value.code == _capi.BO_ERR_INVALID


def test_no_cpu_fallback():
    """Without a device the product path must fail loudly, never fall back to the CPU."""
    import optas_b200
    from optas_b200 import _capi, problems

    if _capi.device_count() > 0:
        pytest.skip("a GPU is present")
    prob = problems.booth()
    with pytest.raises(_capi.BoError) as e:
        optas_b200.B200Solver(prob.opt).setup("ipopt")
    assert e.value.code == _capi.BO_ERR_NO_DEVICE
    solver = optas_b200.B200Solver(prob.opt).setup("ipopt", compile_only=True)
    solver.reset_parameters({"a": 2.0, "b": 7.0})
    with pytest.raises(_capi.BoError) as e:
        solver.solve()
    assert e.value.code == _capi.BO_ERR_NO_DEVICE


def test_product_package_does_not_import_the_oracle():
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, %r); import optas_b200, optas_b200.function, optas_b200.problems; "
            "bad = [m for m in sys.modules if m.split('.')[0] in ('fk_ref','tape_vm','slsqp_driver','kkt_check','hostsim','scipy')]; "
            "print(bad); sys.exit(1 if bad else 0)" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
