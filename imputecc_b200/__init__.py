"""imputecc_b200: B200-native CRWR imputation, a drop-in for ImputeCC's imputation call.

Scope: ImputeMatrix._imputation / random_walk_cpu, and of the rows §8f marks next match_contigs with its
edge weights on the GPU (binning.py).
"""
from .crwr import (DEFAULT_MAX_STEPS, DEFAULT_TOL, Walker, WalkInfo, crwr_impute, imputation_b200, install,
                   markers_first_order, random_walk_b200, shard_rows, stack_row_blocks, symmetrise_normalise,
                   validate_inputs)
from .handoff import contact_thresholds, storage_offenders
from .binning import BinContactWeights, bin_contact_weights, bins_to_arrays, install_binning, match_contigs_b200, precluster_b200

__all__ = ["Walker", "WalkInfo", "crwr_impute", "random_walk_b200", "imputation_b200", "install",
           "symmetrise_normalise", "markers_first_order", "shard_rows", "stack_row_blocks", "validate_inputs",
           "DEFAULT_TOL", "DEFAULT_MAX_STEPS", "BinContactWeights", "bin_contact_weights", "bins_to_arrays", "match_contigs_b200", "precluster_b200", "install_binning", "contact_thresholds", "storage_offenders"]
__version__ = "0.1.0"
