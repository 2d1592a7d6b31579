"""mesh2splat_b200 — B200-native mesh -> 3D gaussian splat conversion path."""
from ._abi import (FLAG_NONE, FLAG_UNCAPPED, LAYOUT_PACKED56, LAYOUT_PLY_COMPRESSED, LAYOUT_PLY_PBR,  # noqa: F401
                   LAYOUT_PLY_STANDARD, LAYOUT_REF96, Primitive, Scene, record_dtype, reference_capacity)
