"""CPU tier: the built library really contains what DESIGN.md describes - the headline kernel
(nx=128, 'std', padding-free) reads and writes tensor memory with 16-column tcgen05 moves,
allocates and frees its columns, and does its arithmetic on the FP64 pipe.  cuobjdump only; no GPU."""
import collections
import os
import re
import shutil
import subprocess

import pytest

from metrotrpl_b200 import _capi

HEADLINE = "trpl_forward_kernelILi4ELi0ELb1"


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")
def test_headline_kernel_sass():
    lib = _capi.library_path()
    assert os.path.exists(lib), "build the library first (__graft_entry__.build())"
    sass = subprocess.check_output(["cuobjdump", "-sass", lib], text=True)
    ops = collections.Counter()
    inside = False
    for line in sass.splitlines():
        if "Function :" in line:
            inside = HEADLINE in line
            continue
        if inside:
            m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
            if m:
                ops[m.group(1)] += 1
    assert ops, "headline instantiation not found in the library"
    fp64 = ops["DFMA"] + ops["DMUL"] + ops["DADD"]
    print({k: v for k, v in ops.most_common(14)})
    # tensor memory as lane-private scratch: wide loads/stores, allocation and release
    assert ops["LDTM.x16"] >= 15 and ops["STTM.x16"] >= 10
    assert ops["UTCATOMSWS.FIND_AND_SET.ALIGN"] >= 1 and ops["UTCATOMSWS.AND"] >= 1
    # the work is FP64 arithmetic (a static count: the cold emission / IRF / ladder code is in the
    # same function; the executed mix is 55% FP64, profiles/r01_ncu_v23_*)
    assert fp64 > 2000 and fp64 > 0.25 * sum(ops.values())
    # shared memory only carries the lane exchange, the late increments and the coefficients
    assert ops["LDS.128"] < 150 and ops["STS.128"] < 100
    # no tensor-core instruction anywhere on this path (nothing is a dense contraction)
    assert not any(k.startswith(("UTCMMA", "UTCHMMA", "HMMA", "DMMA", "QGMMA")) for k in ops)
