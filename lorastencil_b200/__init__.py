"""lorastencil_b200 -- B200-native (sm_100a) LoRAStencil operators.

The product is ``lib/liblorastencil_b200.so`` (hand-written CUDA + C++ host code, C ABI in
``include/lorastencil.h``) and the ``bin/lorastencil_{1d,2d,3d}`` drivers.  This package is the thin
Python host layer over the C ABI:

* ``ops``   -- the reference's host operators (``gpu_box_2d3r`` ... on padded host arrays), same names
  and argument order as ``src/{1d,2d,3d}/*_utils.h`` of zondie17/LoRAStencil;
* ``Plan``  -- device-resident plans on torch CUDA tensors (torch is used for device memory,
  streams and ``torch.distributed`` only);
* ``slab``  -- one-process-per-GPU slab decomposition with halo exchange.

There is no CPU fallback: importing works anywhere, but every compute entry point raises if the
shared library is missing or no CUDA device is present.
"""
from ._lib import (SHAPES, SHAPE_IDS, WEIGHTS_GENERAL, WEIGHTS_REFERENCE, LoraError, build, lib, lib_path,
                   library_built)
from . import ops
from .plan import Plan, decompose_2d, decompose_3d_r2, effective_weights, reference_table

__all__ = ["SHAPES", "SHAPE_IDS", "WEIGHTS_GENERAL", "WEIGHTS_REFERENCE", "LoraError", "build", "lib", "lib_path",
           "library_built", "ops", "Plan", "decompose_2d", "decompose_3d_r2", "effective_weights", "reference_table"]
