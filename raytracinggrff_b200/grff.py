"""GRFF boundary (reference: script/resample_with_ray_tracing.py:74-86, :400-524).

* ``initGET_MW(libname=None)`` — returns the ``PyGET_MW`` callable with exactly the ctypes
  prototype the reference builds at script/resample_with_ray_tracing.py:79-86, bound to
  librtgrff_b200.so (or to any other library exporting ``PyGET_MW``), so
  ``run_ray_tracing_emission(..., grff_lib=<path to librtgrff_b200.so>)`` works unchanged.
* ``get_mw_slice(...)`` — the fastGRFF batched call (script/...:443-446) with the same argument
  list (``tile_pixels`` / ``heap_bytes`` accepted and ignored); numpy arrays in, ``RL_M``
  written in place, status array returned.
"""
from __future__ import annotations

import ctypes

import numpy as np
from numpy.ctypeslib import ndpointer

from . import _lib
from .session import RaySession


def initGET_MW(libname=None):
    _intp = ndpointer(dtype=ctypes.c_int32, flags="F")
    _doublep = ndpointer(dtype=ctypes.c_double, flags="F")
    if libname is None:
        _lib.load()
        libname = str(_lib.LIB_PATH)
    libc_mw = ctypes.CDLL(libname)
    mwfunc = libc_mw.PyGET_MW
    mwfunc.argtypes = [_intp, _doublep, _doublep, _doublep, _doublep, _doublep, _doublep]
    mwfunc.restype = ctypes.c_int
    return mwfunc


def get_mw_slice(Lparms_M, Rparms_M, Parms_M, T_arr, DEM_arr, DDM_arr, RL_M, tile_pixels=None, heap_bytes=None,
                 session=None):
    """fastGRFF.get_mw_slice contract: Lparms_M int32[6] {Npix,Nz,Nf,NT,DEMkey,DDMkey},
    Rparms_M (3,Npix), Parms_M (15,Nz,Npix), RL_M (7,Nf,Npix), all Fortran order.

    The reference calls it with CuPy arrays and reads ``RL_M`` back from the device afterwards
    (script/resample_with_ray_tracing.py:428-452): any array exposing ``__cuda_array_interface__``
    (CuPy, torch, numba) is used where it is and ``RL_M`` is written in place on the device.  With
    numpy arrays the library stages them itself and writes the numpy ``RL_M`` in place.  Returns the
    per-pixel status as a numpy int32 array (0 = ok), which is what the call site tests."""
    ses = session or RaySession(context=_lib.default_context())
    if _lib.cuda_array(RL_M) is None and not isinstance(RL_M, np.ndarray):
        raise TypeError("RL_M must be a numpy array or a device array exposing __cuda_array_interface__ "
                        "(it is written in place)")
    return ses.get_mw_slice(Lparms_M, Rparms_M, Parms_M, RL_M)
