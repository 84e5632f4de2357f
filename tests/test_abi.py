"""CPU: the C-ABI library loads and exports every symbol include/sponet_b200.h declares; run calls fail
loudly without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sponet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sponet_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sponet_b200 import _lib

    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), f"libsponet_b200.so does not export {name}"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert L.sponet_abi_version() == 4


def test_no_cpu_fallback_without_device():
    import torch

    from sponet_b200 import CNVM, CNVMParameters, _lib

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert _lib.device_count() == 0
    p = CNVMParameters(num_opinions=2, num_agents=20, r=1, r_tilde=0.1)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        CNVM(p).simulate(1.0, np.zeros(20, dtype=np.uint8), 3)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under sponet_b200/ may import, link or call it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sponet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "libsponet_oracle" not in text and "sponet_oracle.c" not in text, f


def test_native_kernels_keep_their_state_in_registers():
    """Build sanity: none of the simulation kernels may use local memory (STACK:0).  A spill or an addressable
    copy of the kernel parameters in a hot loop costs a multiple (an out-of-line helper that took the address
    of a parameter member once made the 1-bit CNVM kernels 3x slower without failing any parity test)."""
    import re
    import shutil
    import subprocess

    from sponet_b200 import _lib

    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    kernels = re.findall(r"Function (\S*(?:cnvm_native_kernel|cntm_native_kernel|cnvm_counts_kernel)\S*):\s*\n\s*(.*)", out)
    assert len(kernels) >= 70
    # The 576-thread builds of the COOP kernels (18 warps, 96 registers; template argument 576) keep their per-chain
    # statistics accumulators on the stack -- a few dozen bytes, touched at interval ends only (no local-memory access
    # between the barriers of the event loop in the SASS); everything else must not use local memory at all.
    wide = [(n, u) for n, u in kernels if "Li576E" in n]
    narrow = [(n, u) for n, u in kernels if "Li576E" not in n]
    assert len(wide) >= 8
    bad = [(name, usage) for name, usage in narrow if "STACK:0 " not in usage + " "]
    assert not bad, bad[:3]
    for name, usage in wide:
        stack = int(re.search(r"STACK:(\d+)", usage).group(1))
        assert stack <= 128, (name, usage)
