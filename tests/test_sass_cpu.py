"""The shipped library really contains Blackwell tensor-core / TMA code for the kernels DESIGN.md says use them
(cuobjdump -sass on the in-tree .so; no GPU needed).  UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load,
LDTM = tcgen05.ld, UTCBAR = tcgen05.commit (mnemonics per the B200 profiling guide)."""
import re
import shutil
import subprocess

import pytest

from maf_yolo_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

EXPECT = {
    "gemm_tc_kernel": ["UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"],   # 1x1 / 3x3 s2 implicit GEMM
    "dwpw_kernel": ["UTCHMMA", "UTMALDG", "LDTM", "FFMA2"],       # depth-wise (packed FFMA2) + 1x1 on tcgen05
    "poolpw_kernel": ["UTCHMMA", "UTMALDG", "LDTM"],              # max pool + 1x1
    "stem_conv_kernel": ["UTCHMMA", "LDTM"],                      # software im2col + tcgen05
    "dwconv_kernel": ["FFMA2"],                                   # depth-wise (every instantiation)
    "ELb1E": ["UTMALDG"],                                        # ... its default <K, CB, kTma = true> builds: TMA halo tile
}


@pytest.fixture(scope="module")
def sass_by_kernel():
    if not _lib.LIB_PATH.exists() or not shutil.which(CUOBJDUMP):
        pytest.skip("library or cuobjdump not available")
    out = subprocess.run([CUOBJDUMP, "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in out, "the library was not built for sm_100a"
    kernels = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name is not None:
            kernels[name].append(line)
    return {k: "\n".join(v) for k, v in kernels.items()}


@pytest.mark.parametrize("kernel", sorted(EXPECT))
def test_kernel_uses_the_hardware_it_claims(sass_by_kernel, kernel):
    bodies = [body for name, body in sass_by_kernel.items() if kernel in name and (kernel != "ELb1E" or "dwconv_kernel" in name)]
    assert bodies, f"no {kernel} in the library"
    for mnemonic in EXPECT[kernel]:
        assert all(mnemonic in b for b in bodies), f"{kernel}: {mnemonic} missing in at least one instantiation"
