"""Drop-in for ``raytracingGRFF.gpu_raytrace`` (reference: raytracingGRFF/gpu_raytrace.py).

Same public names, argument meaning, return conventions and error behaviour:

* ``trace_ray(device, ...)``                 — gpu_raytrace.py:414-470
* ``sample_model_with_rays(device, ...)``    — gpu_raytrace.py:712-759
* aliases ``trace_los_dispatch``, ``trace_los_gpu``, ``ray_trace_gpu`` — gpu_raytrace.py:762-780

``device='cuda'`` runs the sm_100a library (librtgrff_b200.so) through ctypes.  There is NO CPU
implementation in this package and no fallback: ``device='cpu'`` raises, and ``fallback_to_cpu``
is accepted for signature compatibility only (a CUDA failure always propagates).  Any other
device string raises the reference's ``ValueError``.

Semantics are those of the reference's CPU path (the parity oracle): float64 ``r_record``,
per-step cross-section ratio, NaN-propagating start — see SURVEY.md §8a "forks".  The cumulative
S variant of the reference's CUDA path is available as ``s_mode='cumulative'``.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from . import _lib
from .session import RaySession

C_R = 2.998e10 / 6.96e10  # build_rays.py:29-32


def _device(device: str) -> str:
    dev = device.lower()
    if dev == "cpu":
        raise RuntimeError("raytracinggrff_b200 is CUDA-only (no CPU path, no fallback); "
                           "use device='cuda', or the reference package for a CPU run.")
    if dev != "cuda":
        raise ValueError(f"Unsupported device '{device}'. Use 'cpu' or 'cuda'.")
    return dev


def _session():
    return RaySession(context=_lib.default_context())


def trace_ray(device, omega_pe_3d, x_grid, y_grid, z_grid, freq_hz, x_start, y_start, z_start, kvec_in_norm,
              dt, n_steps, record_stride=10, trace_crosssections=False, perturb_ratio=1.5, s_mode="per_step"):
    """Trace rays.  Returns ``(r_record, crosssection_record)`` as build_rays.ray_trace does:
    r_record float64 (n_rec, n_rays, 3); crosssection_record a list of n_rec float64 (n_rays,)
    arrays (empty list when cross-sections are not traced)."""
    _device(device)
    ses = _session()
    ses.set_omega_cube(omega_pe_3d, x_grid, y_grid, z_grid, reuse=True)   # same arrays as last call: already there
    mode = {"per_step": _lib.S_PER_STEP, "cumulative": _lib.S_CUMULATIVE}[s_mode]
    r_record, s_record, _ = ses.trace(freq_hz, x_start, y_start, z_start, kvec_in_norm, dt, n_steps,
                                      record_stride, trace_crosssections, perturb_ratio, mode)
    cs = [s_record[i].copy() for i in range(s_record.shape[0])] if trace_crosssections else []
    return r_record, cs


def sample_model_with_rays(device, x_grid, y_grid, z_grid, ne_xyz, te_xyz, b_xyz, r_record, s_arr, ray_start,
                           r_sun_cm, fill_ne=0.0, fill_te=1e4, fill_b=0.0, fallback_to_cpu=True,
                           verbose=True) -> Dict[str, np.ndarray]:
    """Sample model fields along rays.  Returns {ne, te, b, ds, valid_mask, s}, each
    (n_steps, n_rays) (float32 x5, bool), exactly the reference's dict (gpu_raytrace.py:651)."""
    _device(device)
    ses = _session()
    ses.set_field_cubes(x_grid, y_grid, z_grid, ne_xyz, te_xyz, b_xyz, reuse=True)
    return ses.sample(r_record, s_arr, ray_start, r_sun_cm, fill_ne, fill_te, fill_b)


# Backward-compatible aliases (gpu_raytrace.py:762-780).  The *_cpu names are not provided.
def trace_los_gpu_cupy(x_grid, y_grid, z_grid, ne_xyz, te_xyz, b_xyz, r_record, s_arr, ray_start, r_sun_cm,
                       fill_ne=0.0, fill_te=1e4, fill_b=0.0):
    return sample_model_with_rays("cuda", x_grid, y_grid, z_grid, ne_xyz, te_xyz, b_xyz, r_record, s_arr,
                                  ray_start, r_sun_cm, fill_ne, fill_te, fill_b)


def trace_los_dispatch(*args, **kwargs):
    return sample_model_with_rays(*args, **kwargs)


def trace_los_gpu(*args, **kwargs):
    return sample_model_with_rays(*args, **kwargs)


def ray_trace_gpu(omega_pe_3d, x_grid, y_grid, z_grid, freq_hz, x_start, y_start, z_start, kvec_in_norm, dt,
                  n_steps, record_stride=10, trace_crosssections=False, perturb_ratio=2):
    return trace_ray("cuda", omega_pe_3d, x_grid, y_grid, z_grid, freq_hz, x_start, y_start, z_start,
                     kvec_in_norm, dt, n_steps, record_stride, trace_crosssections, perturb_ratio)
