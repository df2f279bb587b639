"""CPU oracle for the bin-contact weights of ImputeCC's ``match_contigs``.  TEST INFRASTRUCTURE ONLY.

Restates the edge-weight loop of /root/reference/Script/utility.py:466-480: for every contig to
bin and every existing bin, the mean imputed contact between the contig and the bin's members,
negated,

    hic_sum = 0
    for j in range(len(bins[b])): hic_sum += E[contig, bins[b][j]]      # scalar look-ups in a float32 LIL matrix
    weight = -hic_sum / len(bins[b])

The look-ups return numpy float32 scalars, so the sum runs in float32, member after member in the
order of the bin's list, and the weight of a bin without any contact is -0.0.  Only ``tests/`` and
measurement scripts may import this module.  Pinned against the unmodified reference function by
``oracle/make_golden_binweights.py`` -> ``tests/golden/binweights_*.npz`` (``tests/test_binweights.py``).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def bin_contact_weights(E, to_bin, bin_ptr, bin_members) -> np.ndarray:
    """weights[r, b] for contigs ``to_bin[r]`` and bins ``bin_members[bin_ptr[b]:bin_ptr[b+1]]`` (local indices)."""
    E = sp.csr_matrix(E)
    dtype = E.dtype.type
    n_bins = len(bin_ptr) - 1
    out = np.empty((len(to_bin), n_bins), dtype=E.dtype)
    for r, c in enumerate(to_bin):
        row = np.asarray(E.getrow(int(c)).todense()).ravel()                 # E[c, :] as the LIL look-ups see it
        for b in range(n_bins):
            members = bin_members[bin_ptr[b]:bin_ptr[b + 1]]
            vals = row[members]
            # float32 running sum in list order (np.cumsum accumulates sequentially); `0 + x` of the first add is exact
            s = np.cumsum(vals, dtype=E.dtype)[-1] if len(vals) else dtype(0)
            out[r, b] = -s / dtype(len(members))
    return out


def bin_contact_weights_loops(E, to_bin, bin_ptr, bin_members) -> np.ndarray:
    """The same with the reference's own statements (scalar LIL look-ups, Python accumulation); small cases only."""
    lil = sp.lil_matrix(E)
    n_bins = len(bin_ptr) - 1
    out = np.empty((len(to_bin), n_bins), dtype=lil.dtype)
    for r, c in enumerate(to_bin):
        for b in range(n_bins):
            hic_sum = 0
            n_contigs = int(bin_ptr[b + 1] - bin_ptr[b])
            for j in range(n_contigs):
                hic_sum += lil[int(c), int(bin_members[bin_ptr[b] + j])]     # utility.py:478
            out[r, b] = -hic_sum / n_contigs                                 # utility.py:480
    return out
