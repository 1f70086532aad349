"""Mirror of ``eradiate.kernel`` (``src/eradiate/kernel/__init__.pyi``)."""

from ._bitmap import Bitmap
from ._kernel_dict import (
    KernelContext,
    KernelSceneParameterFlags,
    KernelSceneParameterMap,
    SceneParameter,
    SearchSceneParameter,
)
from ._render import (
    DeviceScene,
    MitsubaObjectWrapper,
    SceneParameters,
    SeedState,
    develop,
    get_seed_state,
    mi_load_dict,
    mi_render,
    mi_traverse,
    render,
)
from ._scene import BSDF, Emitter, Medium, PhaseFunction, Scene, Sensor, Shape
from ._types import ScalarTransform4f, VolumeGrid, map_cube, map_unit_cube

__all__ = [
    "BSDF",
    "Bitmap",
    "DeviceScene",
    "Emitter",
    "KernelContext",
    "KernelSceneParameterFlags",
    "KernelSceneParameterMap",
    "Medium",
    "MitsubaObjectWrapper",
    "PhaseFunction",
    "ScalarTransform4f",
    "Scene",
    "SceneParameter",
    "SceneParameters",
    "SearchSceneParameter",
    "SeedState",
    "Sensor",
    "Shape",
    "VolumeGrid",
    "develop",
    "get_seed_state",
    "map_cube",
    "map_unit_cube",
    "mi_load_dict",
    "mi_render",
    "mi_traverse",
    "render",
]
