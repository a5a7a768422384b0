"""Split of the FEP pair list over ranks (one rank = one GPU).

Contiguous ranges of i-entries with about equal pair counts, found with the rule the reference
uses to balance its FEP lists over OpenMP threads (src/gromacs/nbnxm/pairlist.cpp:2786-2838):
walk the entries once and move on to the next destination when adding the entry would overshoot
the per-destination target by more than the current shortfall.  libfepb200 applies the same rule
inside fepb200_set_list(); this module exists so that host code and tests can predict the shards.
"""
from __future__ import annotations

import numpy as np


def balanced_ranges(jindex, nranks: int) -> list[tuple[int, int]]:
    jindex = np.asarray(jindex, np.int64)
    nri = jindex.shape[0] - 1
    first = [nri] * (nranks + 1)
    first[0] = 0
    total = int(jindex[-1]) if nri > 0 else 0
    target = (total + nranks - 1) // nranks
    dest, have = 0, 0
    for n in range(nri):
        nrj = int(jindex[n + 1] - jindex[n])
        if dest + 1 < nranks and have > 0 and have + nrj - target > target - have:
            dest += 1
            first[dest] = n
            have = 0
        have += nrj
    return [(first[r], first[r + 1]) for r in range(nranks)]
