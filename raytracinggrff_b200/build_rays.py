"""Drop-in for the hot-path part of ``raytracingGRFF.build_rays`` (reference:
raytracingGRFF/build_rays.py:128-248): ``ray_trace`` with the same signature and return
convention, executed by the sm_100a library.  The MAS/psipy resampling helpers of that module
(build_rays.py:35-125, :251-395) are out of scope."""
from __future__ import annotations

from .gpu_raytrace import C_R, trace_ray

__all__ = ["C_R", "ray_trace"]


def ray_trace(omega_pe_3d, x_grid, y_grid, z_grid, freq_hz, x_start, y_start, z_start, kvec_in_norm, dt,
              n_steps, record_stride=10, trace_crosssections=False, cross_section_stride=1, perturb_ratio=2):
    """build_rays.py:128-130.  ``cross_section_stride`` is accepted and ignored, as in the reference."""
    return trace_ray("cuda", omega_pe_3d, x_grid, y_grid, z_grid, freq_hz, x_start, y_start, z_start,
                     kvec_in_norm, dt, n_steps, record_stride=record_stride,
                     trace_crosssections=trace_crosssections, perturb_ratio=perturb_ratio)
