#!/usr/bin/env python
"""Generate golden OUTPUT vectors by running the UNMODIFIED reference (/root/reference) in the
build container.  The reference is Python and cannot travel to the GPU box, so its outputs on
the seeded cases of tests/cases.py are committed here as small .npz fixtures; inputs are
regenerated from code.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference imports matplotlib at module top (raytracingGRFF/build_rays.py:15-17); matplotlib
is not installed here, so a two-line stub package is put on sys.path first.  Nothing else of the
reference is touched, and no reference source is copied into this repo.
"""
from __future__ import annotations

import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
REF = Path("/root/reference")


def _import_reference():
    stub = Path(tempfile.mkdtemp(prefix="mplstub_"))
    (stub / "matplotlib").mkdir()
    (stub / "matplotlib" / "__init__.py").write_text("def use(*a, **k):\n    pass\n")
    (stub / "matplotlib" / "pyplot.py").write_text("")
    sys.path.insert(0, str(stub))
    sys.path.insert(0, str(REF))
    from raytracingGRFF.build_rays import ray_trace            # noqa: E402
    from raytracingGRFF.gpu_raytrace import sample_model_with_rays  # noqa: E402
    return ray_trace, sample_model_with_rays


def main():
    import cases
    from raytracinggrff_b200 import synthetic

    ray_trace, sample_model_with_rays = _import_reference()
    out = Path(__file__).resolve().parent

    for name in cases.TRACE_CASES:
        kw = cases.trace_case(name)
        t0 = time.time()
        r_record, cs = ray_trace(**kw)
        print(f"{name}: reference ray_trace {time.time() - t0:.1f}s r_record {r_record.shape}")
        np.savez_compressed(out / f"trace_{name}.npz", r_record=r_record,
                            s_record=np.array(cs) if len(cs) else np.zeros((0,)))

    for seed in (1, 2, 3):
        xg, yg, zg, ne, te, b, r_record, s_arr, ray_start = cases.sampler_fixture(seed)
        o = sample_model_with_rays("cpu", xg, yg, zg, ne, te, b, r_record, s_arr, ray_start, r_sun_cm=1.0)
        np.savez_compressed(out / f"sampler_fixture_seed{seed}.npz", **o)

    # config-1 shape at reduced size, real r_sun_cm
    xg, yg, zg, ne, te, b, r_record, s_arr, ray_start = cases.los_sampler_case(24, 48, 40, seed=0)
    o = sample_model_with_rays("cpu", xg, yg, zg, ne, te, b, r_record, s_arr, ray_start, r_sun_cm=6.957e10)
    np.savez_compressed(out / "sampler_c1_small.npz", **o)

    # sampler on real traced paths (float64 r_record, NaN S after exit), config-3 physics
    kw = cases.trace_case("corona_cs")
    g = np.load(out / "trace_corona_cs.npz")
    c = synthetic.corona_cube(48, 3.0)
    ray_start = np.column_stack([kw["x_start"], kw["y_start"], kw["z_start"]])
    o = sample_model_with_rays("cpu", c["x_grid"], c["y_grid"], c["z_grid"], c["ne"], c["te"], c["b"],
                               g["r_record"], g["s_record"], ray_start, r_sun_cm=6.957e10)
    np.savez_compressed(out / "sampler_on_traced_paths.npz", **o)
    # per-frequency presets of the publication driver (script/pub/TbSpectra_gen.py:27-88).  The module
    # imports matplotlib/astropy/psipy at top level, so only the four pure functions are extracted
    # from its source and executed.
    import ast
    import json
    src = (REF / "script" / "pub" / "TbSpectra_gen.py").read_text()
    want = {"_lowband_params", "_interp_log_freq_params", "_highband_params", "select_params"}
    code = "\n\n".join(ast.get_source_segment(src, n) for n in ast.parse(src).body
                       if isinstance(n, ast.FunctionDef) and n.name in want)
    ns = {"np": np}
    exec(code, ns)
    freqs = [20e6, 30e6, 75e6, 100e6, 150e6, 150.0001e6, 200e6, 279e6, 280e6, 400e6, 550e6, 700e6, 800e6, 1.2e9]
    json.dump({repr(f): ns["select_params"](f) for f in freqs}, open(out / "select_params.json", "w"), indent=1)
    # image-plane operations: the reference's own patch_nan_emission_map (util.py is loaded by path: the
    # package __init__ would import matplotlib/psipy users) and the scipy call the workflow makes
    import importlib.util
    from scipy.ndimage import gaussian_filter
    spec = importlib.util.spec_from_file_location("ref_util", REF / "raytracingGRFF" / "util.py")
    ref_util = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_util)
    img = {}
    for name in cases.IMAGE_CASES:
        a = cases.image_case(name)
        img[f"patch_{name}"] = ref_util.patch_nan_emission_map(a)
    smooth = ref_util.patch_nan_emission_map(cases.image_case("sparse"))
    for sg in cases.BEAM_SIGMAS:
        img[f"beam_{sg}"] = gaussian_filter(smooth, sigma=sg)
    img["beam_nan_2.5"] = gaussian_filter(cases.image_case("sparse"), sigma=2.5)
    np.savez_compressed(out / "image_ops.npz", **img)
    print("golden vectors written to", out)


if __name__ == "__main__":
    main()
