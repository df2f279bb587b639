"""Hand-off of the imputed matrix to the clustering stages.

Rank 2 - the LIL hand-off (/root/reference/Script/imputation.py:129, read back at utility.py:477).  ``.tolil()`` turns
the m x m result into Python lists of Python floats: 21 ms at 8,000 markers, 111 ms at 32,000 (up to 430 ms for a dense
imputation) - ten times the whole GPU imputation.  Every consumer of ``imputed_matrix`` only calls ``.tocoo()``,
``.tolil()`` and ``.copy().tocsr()`` on it (pre_clustering.py:47-61, final_cluster.py:334, 393, 406-407), which a CSR
matrix offers too, and :func:`imputecc_b200.match_contigs_b200` takes the CSR form directly: ``hand_off="csr"`` of
:func:`imputecc_b200.imputation_b200` / :func:`imputecc_b200.install` skips the conversion.

Rank 4 - the on-disk hand-off (ImputeCC.py:241, 338: the whole ``ImputeMatrix`` object pickled into
``ImputeCC_storage.gz`` by utility.py:26-46 and unpickled by another process).  The object must therefore hold plain
numpy / scipy / builtin members only - :func:`storage_offenders` checks that - and the two contact thresholds the
clustering derives from it, recomputed from ``imputed_matrix.tocoo().data`` inside the decontamination loop
(final_cluster.py:393, 406-407) and in PreCluster (pre_clustering.py:60-61), are computed once by
:func:`contact_thresholds`.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def contact_thresholds(imputed_matrix, w_intra, w_inter):
    """``(np.percentile(data, w_intra), np.percentile(data, w_inter))`` of the stored values of the imputed matrix -
    what pre_clustering.py:60-61 and final_cluster.py:406-407 compute (twice per call, through a COO conversion) - or
    ``None`` for an empty matrix, the case both callers branch on first (pre_clustering.py:47, final_cluster.py:393).
    The values do not depend on the sparse format: a percentile is a function of the multiset of stored values."""
    data = imputed_matrix.data if isinstance(imputed_matrix, (sp.csr_matrix, sp.csc_matrix, sp.coo_matrix)) \
        else imputed_matrix.tocoo().data
    if isinstance(imputed_matrix, (sp.csr_matrix, sp.csc_matrix)) and not imputed_matrix.has_canonical_format:
        data = imputed_matrix.tocoo().data
    if len(data) == 0:
        return None
    # (one call per threshold, scalar q, exactly as the reference: numpy then keeps the data's own float type)
    return np.percentile(data, w_intra), np.percentile(data, w_inter)


_PLAIN = (type(None), bool, int, float, complex, str, bytes, np.generic, np.ndarray, sp.spmatrix)


def storage_offenders(obj, _path="obj", _seen=None):
    """Names of members of ``obj`` (recursively through dicts, lists, tuples, sets and instance attributes) that are not
    plain builtin / numpy / scipy.sparse values - e.g. a torch tensor, a ctypes handle, a CUDA stream - and would tie
    ``ImputeCC_storage.gz`` to this package or to a GPU.  Empty list = safe to pickle for the reference's loader."""
    seen = _seen if _seen is not None else set()
    if id(obj) in seen:
        return []
    seen.add(id(obj))
    if isinstance(obj, _PLAIN) or sp.issparse(obj):
        return []
    if isinstance(obj, dict):
        out = []
        for k, v in obj.items():
            out += storage_offenders(k, "%s key %r" % (_path, k), seen) + storage_offenders(v, "%s[%r]" % (_path, k), seen)
        return out
    if isinstance(obj, (list, tuple, set, frozenset)):
        out = []
        for i, v in enumerate(obj):
            out += storage_offenders(v, "%s[%d]" % (_path, i), seen)
        return out
    mod = type(obj).__module__ or ""
    if mod.split(".")[0] in ("torch", "ctypes", "imputecc_b200", "cuda"):
        return ["%s: %s.%s" % (_path, mod, type(obj).__name__)]
    if hasattr(obj, "__dict__"):
        out = []
        for k, v in vars(obj).items():
            out += storage_offenders(v, "%s.%s" % (_path, k), seen)
        return out
    return []
