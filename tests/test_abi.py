"""CPU: the C-ABI library loads, exports every symbol include/crowdnav_b200.h declares, and the ctypes mirror
(crowdnav_dsrnn_b200/abi.py) has the same struct layout as the C header (checked with a gcc probe)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from crowdnav_dsrnn_b200 import _lib, abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "crowdnav_b200.h")


def declared_functions():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"\b(cn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_functions()
    assert len(names) >= 20
    assert set(names) == set(_lib.SYMBOLS), "python binding table and header disagree"
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), n
    assert lib.cn_abi_version() == abi.ABI_VERSION


def test_struct_layout_matches_header():
    probe = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "crowdnav_b200.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu ", sizeof(CnConfig), sizeof(CnStateView), sizeof(CnObsOut), sizeof(CnStepOut),
               sizeof(CnDsrnnWeights), sizeof(CnDsrnnIO));
        printf("%zu %zu %zu %zu %zu %d\n", offsetof(CnConfig, scenarios), offsetof(CnConfig, goal_change_steps),
               offsetof(CnConfig, base_seed), offsetof(CnConfig, time_step), offsetof(CnConfig, orca_neighbor_dist), CN_INFO_DIM);
        return 0;
    }'''
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "p.c"), os.path.join(d, "p")
        open(src, "w").write(probe)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        got = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [C.sizeof(abi.CnConfig), C.sizeof(abi.CnStateView), C.sizeof(abi.CnObsOut), C.sizeof(abi.CnStepOut),
            C.sizeof(abi.CnDsrnnWeights), C.sizeof(abi.CnDsrnnIO), abi.CnConfig.scenarios.offset,
            abi.CnConfig.goal_change_steps.offset, abi.CnConfig.base_seed.offset, abi.CnConfig.time_step.offset,
            abi.CnConfig.orca_neighbor_dist.offset, abi.INFO_DIM]
    assert got == want


def test_argument_errors_are_reported_not_crashed():
    lib = _lib.load()
    cfg = abi.CnConfig()          # abi_version 0: rejected
    assert lib.cn_env_state_bytes(C.byref(cfg), 4) == 0
    out = C.c_void_p()
    rc = lib.cn_env_create(C.byref(cfg), 4, 0, None, 0, C.byref(out))
    assert rc == -1 and b"abi_version" in lib.cn_last_error()
    assert lib.cn_dsrnn_workspace_bytes(0, 5) == 0 and lib.cn_dsrnn_workspace_bytes(16, 5) > 0


def test_oracle_is_not_imported_by_the_product():
    """The product package must never import / call anything under oracle/."""
    pkg = os.path.join(ROOT, "crowdnav_dsrnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text.replace(
                    "oracle/crowd_oracle.c restates", "").replace("oracle/orca_core.h", ""), os.path.join(dirpath, f)
    code = "import sys; import crowdnav_dsrnn_b200.envs, crowdnav_dsrnn_b200.model, crowdnav_dsrnn_b200.crowd_sim_dict; " \
           "assert not [m for m in sys.modules if m.split('.')[0] == 'oracle']"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def test_graft_entry_build_runs():
    """The driver's "does it build" check: __graft_entry__.build() must make + load the library and the C oracle."""
    sys.path.insert(0, ROOT)
    import __graft_entry__

    pkg = __graft_entry__.build()
    assert pkg.__name__ == "crowdnav_dsrnn_b200"
    assert os.path.exists(_lib.LIB_PATH)
