"""raytracinggrff_b200 — B200 (sm_100a) implementation of the per-ray hot path of
peijin94/raytracingGRFF behind the reference's own Python API.

Exports mirror ``raytracingGRFF/__init__.py:3-15`` for the hot path (``C_R``, ``ray_trace``,
``trace_ray``, ``sample_model_with_rays``, ``patch_nan_emission_map``, ``resample_to_xyz_cube``); reading
MAS HDF files needs psipy and is out of scope: ``load_mas_var_filtered`` is exported for
import compatibility and serves variables of in-memory models only.  Nothing here imports ``oracle/``.
"""
from .build_rays import ray_trace
from .cubes import load_mas_var_filtered, resample_to_xyz_cube
from .gpu_raytrace import C_R, sample_model_with_rays, trace_ray
from .grff import get_mw_slice, initGET_MW
from .session import RaySession
from .util import patch_nan_emission_map

__all__ = ["C_R", "RaySession", "get_mw_slice", "initGET_MW", "load_mas_var_filtered", "patch_nan_emission_map",
           "ray_trace", "resample_to_xyz_cube", "sample_model_with_rays", "trace_ray"]
