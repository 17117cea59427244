"""csp_b200 -- B200 (sm_100a) accelerator for the data-parallel hot path of CoordinateSplittingPTrees.jl.

Host side (tree growth, the reference's API names): tree.py / api.py.  Device side (hand-written CUDA behind the C ABI of
include/csp_b200.h): device.py.  Synthetic benchmark configurations: synthetic.py.
"""
from ._native import (ArgumentError, BoundsError, CspError, DimensionMismatch, ErrorException, NoDeviceError, UndefRefError)
from .tree import (Box, World, addpoint, addpoints, boxbounds, collect, degree, getleaf, getroot, isfake, isleaf, isroot, leaf_id,
                   leaves, length, meta, ndims, node_id, position, quadnoise_objective, rosenbrock_objective, split, value)
from .device import DeviceTree, LeafModels, device_count, kernel_launches, load_blob, pack_blob, save_blob
from .api import (SymmetricArray, coefficients_p, coefficients_p_sparse, device_tree, extrema, find_leaf_at, fit_leaves, fit_quadratic, fit_quadratic_gdiag, fit_quadratic_native,
                  locate_eval, maximum, minimum, modelvalue, modelvalue_native, polynomial_full, polynomial_minimal, possemidef)
from . import shard, synthetic
