"""Host-side model check of conv3x3_dx_kernel's mbarrier protocol (tools/sim_dx_protocol.py): random interleavings of
the producer, the two UMMA-issuing warps and the epilogue groups, with asynchronous TMA / tcgen05.commit completions and
the hardware's parity semantics of mbarrier waits.  Every configuration the host code can choose (launch geometry in
dd_launch_conv3x3_dx: sub-tiles per box imply one box per unit) must neither deadlock nor read a stage / accumulator in
the wrong fill -- the failure mode of sharing one stage ring between the two issuing warps (they lap each other)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import sim_dx_protocol as sim  # noqa: E402


@pytest.mark.parametrize("mma_warps", [1, 2])
def test_protocol_has_no_deadlock_or_stale_read(mma_warps):
    runs = 0
    for a_stages in (2, 3, 4, 5, 6):
        if mma_warps == 2 and a_stages < 4:
            continue                                   # host: two issuing warps need a ring of >= 2 stages each
        for kchunks in (1, 2, 3, 5):
            for nsub, nbuf, ngroups in ((1, 4, 4), (2, 4, 4), (1, 2, 2)):
                if nsub == 2 and kchunks != 1:
                    continue                           # two sub-tiles per box <=> Cin/g == 32 <=> one box per unit
                for res in (False, True):
                    for units, m_tiles in ((20, 7), (9, 3), (4, 12), (13, 1)):
                        for seed in range(4):
                            r, errs = sim.simulate(seed, units, m_tiles, a_stages=a_stages, kchunks=kchunks, nsub=nsub,
                                                   nbuf=nbuf, ngroups=ngroups, res=res, mma_warps=mma_warps)
                            assert r == "ok" and not errs, (a_stages, kchunks, nsub, nbuf, res, units, m_tiles, seed, r, errs[:2])
                            runs += 1
    assert runs > 500
