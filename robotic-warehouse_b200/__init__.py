"""rware_b200 -- B200-native batched Multi-Robot Warehouse (RWARE) step path.

Directory name: `robotic-warehouse_b200/` (import it as `rware_b200` through the shim module at the repo root).
Drop-in for the step path of semitable/robotic-warehouse: `Warehouse` (Gymnasium surface, one world),
`VecWarehouse` (E worlds, tensors), and the registered `rware-*` ids.
"""
from .layout import layout_from_params, layout_from_str  # noqa: F401
from .registration import env_kwargs, full_registration, image_registration, make, make_vec, registry  # noqa: F401
from .vec import VecWarehouse  # noqa: F401
from . import wrappers  # noqa: F401
from .warehouse import (Action, Agent, Direction, ImageLayer, ObservationType, RewardType, Shelf,  # noqa: F401
                        Warehouse)

__version__ = "0.1.0"
