"""Drop-in for the `spconv` v1.0 Python surface that PCDet uses (surface B2):

    import pcdet_b200.spconv as spconv          # or: sys.modules['spconv'] = pcdet_b200.spconv

SparseConvTensor / SparseModule / SparseSequential / SubMConv3d / SparseConv3d / SparseInverseConv3d,
spconv.utils.VoxelGenerator and spconv.ops.get_indice_pairs keep their names, arguments, parameter
layout and state-dict keys; the arithmetic runs in libpcdet_b200.so on sm_100a.
"""
from . import ops, utils  # noqa: F401
from .graph import GraphedSparseModule  # noqa: F401
from .conv import SparseConv3d, SparseConvolution, SparseInverseConv3d, SubMConv3d  # noqa: F401
from .modules import SparseModule, SparseSequential  # noqa: F401
from .pool import SparseMaxPool3d  # noqa: F401
from .tensor import SparseConvTensor  # noqa: F401


def install_as_spconv():
    """Registers this package under the name `spconv` so unmodified PCDet files import it."""
    import sys
    sys.modules["spconv"] = sys.modules[__name__]
    sys.modules["spconv.utils"] = utils
    sys.modules["spconv.ops"] = ops
