"""Regression guard for the register-stage kernel's code generation ("ptxas uniformity").

ptxas keeps the step loop of qls_rs_pass_kernel on the uniform datapath (gate matrices as uniform-register
operands of DFMA / FFMA, pass description through LDCU) only while it can prove the loop warp-uniform; small
source changes make it fall back to per-thread LDC operands for the WHOLE loop, which costs ~25 % of the
headline without failing any parity test.  This test reads the SASS of the object file the build produced."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "qc-quantum-linear-systems_b200", "csrc", "rs_pass.o")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not available")
def test_rs_pass_keeps_uniform_register_operands():
    if not os.path.exists(OBJ):
        import qls_b200.build as build

        build.build()
    sass = subprocess.run([CUOBJDUMP, "-sass", OBJ], capture_output=True, text=True, check=True, timeout=300).stdout
    per_kernel = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per_kernel[name] = {"DFMA": 0, "DFMA_UR": 0, "FFMA": 0, "FFMA_UR": 0, "local": 0}
        elif name:
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if not m:
                continue
            op = m.group(1)
            if op in ("DFMA", "FFMA"):
                per_kernel[name][op] += 1
                per_kernel[name][op + "_UR"] += bool(re.search(r"\bUR\d+\b", line))
            per_kernel[name]["local"] += op in ("STL", "LDL")
    kernels = {k: v for k, v in per_kernel.items() if "qls_rs_pass_kernel" in k}
    assert len(kernels) == 4, sorted(per_kernel)
    for k, c in kernels.items():
        fp64 = "kernelId" in k
        total, ur = (c["DFMA"], c["DFMA_UR"]) if fp64 else (c["FFMA"], c["FFMA_UR"])
        assert total > 4000, (k, c)
        assert ur >= 0.94 * total, f"{k}: only {ur} of {total} FMAs take a uniform-register operand"
        assert c["local"] == 0, f"{k}: spills ({c['local']} local-memory instructions)"
