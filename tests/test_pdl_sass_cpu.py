"""Structural check of programmatic dependent launch on the built library's SASS (no GPU needed): EVERY kernel contains
griddepcontrol.wait (ACQBULK) and griddepcontrol.launch_dependents (PREEXIT), and no instruction that touches global memory --
loads, stores, atomics / reductions, TMA loads / stores / prefetches, bulk copies -- precedes the wait in program order.  That is
the property the scheme rests on (common.cuh: launch_kernel / pdl_sync): a kernel launched with programmatic stream serialization
may start before its predecessor has finished, so it must not see or change global memory before the wait."""
import os
import re
import shutil
import subprocess

import pytest

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

# global-memory instructions of sm_100 SASS (shared-memory / mbarrier forms such as LDS, STS, SYNCS.* are not in the list, nor is the
# L1 invalidate CCTL.IVALL that belongs to the acquire side of a cluster barrier)
GLOBAL_OP = re.compile(r"^(@!?U?P\d+\s+)?(LDG|STG|LD|ST|ATOM|ATOMG|RED|REDG|LDGSTS|UTMALDG|UTMASTG|UTMAPF|UTMAREDG|UBLKCP|UBLKRED|UBLKPF)\b")


@pytest.fixture(scope="module")
def sass():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump (CUDA toolkit) not found")
    from molclr_b200.build import build
    return subprocess.run([CUOBJDUMP, "-sass", build()], capture_output=True, text=True, timeout=600).stdout


def test_every_kernel_waits_before_its_first_global_access(sass):
    funcs, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
        if cur is not None and m:
            cur.append(m.group(1).strip())
    assert len(funcs) >= 100, len(funcs)
    problems = []
    for name, ins in funcs.items():
        waits = [i for i, t in enumerate(ins) if "ACQBULK" in t]
        if not waits or not any("PREEXIT" in t for t in ins):
            problems.append((name, "no griddepcontrol.wait / launch_dependents"))
            continue
        early = [t for t in ins[:waits[0]] if GLOBAL_OP.match(t)]
        if early:
            problems.append((name, "global access before the wait: " + early[0]))
    assert not problems, problems[:5]


def test_tensor_core_and_tma_instructions_are_the_blackwell_ones(sass):
    """The contraction and tile-movement kernels are written against tcgen05 / TMEM / TMA: the library's SASS holds UTCHMMA (incl. the
    CTA-pair form), TMEM loads, TMA tensor loads and stores and bulk copies -- and no legacy warp-level MMA."""
    count = lambda pat: len(re.findall(pat, sass))
    assert count(r"\bUTCHMMA\b") > 100 and count(r"\bUTCHMMA\.2CTA\b") > 100
    assert count(r"\bLDTM\b") > 50 and count(r"\bUTMALDG\.") > 300 and count(r"\bUTMASTG\.") > 10 and count(r"\bUBLKCP\b") > 5
    assert count(r"\bHMMA\b") == 0 and count(r"\bHGMMA\b") == 0
