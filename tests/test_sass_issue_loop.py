"""Static check of the shipped library's SASS (no GPU): the tcgen05.mma issue loops of the product kernels must be warp-uniform.

An issue loop inside `if (lane == 0)` makes ptxas wrap every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop that moves the
descriptors into uniform registers - ~100 cycles per MMA, the "issue floor" of round 2 (profiles/r02u_pair_roles.md).  The test fails if
that pattern comes back next to an MMA of the conv / matching kernels.  Allowed: the 12 MMAs of the retired second-issuer block of the
64 -> 64 instance (never executed) and the opt-in operand-swapped experiment kernel."""
import collections
import functools
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "qmri-pnp-recon-poc_b200", "lib", "libqmri_b200.so")


@functools.lru_cache(maxsize=1)
def _sass_by_kernel():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, lines = None, collections.defaultdict(list)
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", l)
        if m and cur:
            lines[cur].append(m.group(1))
    return lines


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
def test_mma_issue_loops_are_warp_uniform():
    kernels = _sass_by_kernel()
    seen = collections.Counter()
    for name, ins in kernels.items():
        fam = next((f for f in ("tc_conv3x3_pair_kernel", "tc_conv_kernel", "match_tc_kernel") if f in name), None)
        if fam is None:
            continue
        mma = [i for i, x in enumerate(ins) if "UTCHMMA" in x]
        assert mma, f"{name}: no tcgen05.mma in a tensor-core kernel"
        waterfall = sum(1 for i in mma if any("R2UR.BROADCAST" in x or "BRA.U.ANY" in x for x in ins[max(0, i - 10):i + 2]))
        allowed = 12 if "tc_conv3x3_pair_kernelILi128ELi1E" in name else 0
        assert waterfall <= allowed, f"{name}: {waterfall} of {len(mma)} UTCHMMA sit in a register-to-uniform waterfall loop (issue loop not warp-uniform)"
        if fam == "tc_conv3x3_pair_kernel":
            assert any("UTCHMMA.2CTA" in x for x in ins), f"{name}: no cta_group::2 MMA"
        seen[fam] += 1
    assert seen["tc_conv3x3_pair_kernel"] == 3 and seen["tc_conv_kernel"] == 6 and seen["match_tc_kernel"] >= 20, seen


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
def test_64_channel_instance_uses_bulk_tensor_stores_and_loads():
    """The 64 -> 64 conv instance sends its tiles out with UTMASTG (bulk tensor stores) and has no local-memory spills."""
    kernels = _sass_by_kernel()
    name = next(n for n in kernels if "tc_conv3x3_pair_kernelILi128ELi1E" in n)
    ins = kernels[name]
    assert sum("UTMASTG" in x for x in ins) >= 2
    assert sum("UTMALDG" in x for x in ins) >= 8
    assert not any(x.startswith(("LDL", "STL")) for x in ins), "register spills in the 64 -> 64 conv kernel"
