"""B200-native sweeping MPS classifier: drop-in for the hot path of francescovidaich964/TensorNetworkForML.

Public surface = the reference's modules, under the same names:
    Network_class.Network, Tensor_class.Tensor, custom_linalg_tools.contract, data_generator.*
The module names are also registered at top level (``import Network_class``) unless a module of that name is
already imported, so ``.dat`` pickles written by the reference load here and vice versa.
"""
import sys as _sys

from . import Tensor_class, custom_linalg_tools, data_generator, Network_class  # noqa: F401
from .Network_class import Network  # noqa: F401
from .Tensor_class import Tensor  # noqa: F401
from .custom_linalg_tools import contract  # noqa: F401

for _name, _mod in (("Tensor_class", Tensor_class), ("custom_linalg_tools", custom_linalg_tools),
                    ("data_generator", data_generator), ("Network_class", Network_class)):
    _sys.modules.setdefault(_name, _mod)

__all__ = ["Network", "Tensor", "contract", "Network_class", "Tensor_class", "custom_linalg_tools", "data_generator"]
