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


TEAM = "trpl_team_forward_kernelILi4ELi1ELb1"


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")
def test_two_warp_team_kernel_sass():
    """The configs[3] instantiation (traps, nx = 256) is the two-warp team: named barriers of 64
    threads, tensor-memory moves like the headline kernel, FP64 arithmetic, and next to no spills
    (the one-warp kernel with 8 nodes per lane it replaces had ~900 LDL/STL in its SASS)."""
    lib = _capi.library_path()
    sass = subprocess.check_output(["cuobjdump", "-sass", lib], text=True)
    ops = collections.Counter()
    named_bar = 0
    inside = False
    for line in sass.splitlines():
        if "Function :" in line:
            inside = TEAM in line
            continue
        if inside:
            m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)(.*)", line)
            if m:
                ops[m.group(1)] += 1
                if m.group(1).startswith("BAR.SYNC") and "0x40" in m.group(2):
                    named_bar += 1
    assert ops, "team instantiation not found in the library"
    print({k: v for k, v in ops.most_common(14)}, named_bar)
    assert named_bar >= 30                                     # bar.sync <1+team>, 64
    assert ops["LDTM.x16"] >= 10 and ops["STTM.x16"] >= 8
    fp64 = ops["DFMA"] + ops["DMUL"] + ops["DADD"]
    assert fp64 > 2000
    assert ops["LDL"] + ops["STL"] + ops["LDL.64"] + ops["STL.64"] + ops["LDL.128"] + ops["STL.128"] < 250
    assert not any(k.startswith(("UTCMMA", "UTCHMMA", "HMMA", "DMMA", "QGMMA")) for k in ops)

