"""A host written in plain C against include/rtgrff.h (examples/c_host_map.c): the header is C99-clean, the
library links without Python or torch, fails loudly without a GPU, and — on the B200 — renders the same map as the
Python API does on the same model."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
EXE = ROOT / "build" / "c_host_map"


def _build():
    from raytracinggrff_b200.build import build_library
    so = build_library()
    EXE.parent.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}",
                    str(ROOT / "examples" / "c_host_map.c"), "-o", str(EXE), f"-L{so.parent}", "-lrtgrff_b200",
                    f"-Wl,-rpath,{so.parent}", "-lm"], check=True)


def test_c_host_compiles_links_and_refuses_to_run_without_a_gpu():
    _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([str(EXE), "8", "24", "1e8"], capture_output=True, text=True)
    assert p.returncode == 3 and "no CUDA device" in p.stderr and p.stdout == ""


@pytest.mark.gpu
def test_c_host_renders_the_same_map_as_the_python_api(session):
    _build()
    n_pix, n, freq = 16, 48, 120e6
    p = subprocess.run([str(EXE), str(n_pix), str(n), repr(freq)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    lines = p.stdout.splitlines()
    assert lines[0].startswith("# rtgrff_b200")
    img = np.array([float(v) for v in lines[1:]]).reshape(2, n_pix, n_pix)
    # the same model and rays through the Python API
    extent, x_fov, z_obs = 3.0, 1.3, 3.0
    g = -extent + 2.0 * extent * np.arange(n) / (n - 1)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    r = np.sqrt(X * X + Y * Y + Z * Z)
    inside = r < 0.999999
    rs = np.where(inside, 1.0, r)
    ne = np.where(inside, 0.0, 4.2e4 * 10.0 ** (4.32 / rs))
    te = np.where(inside, 1.0e4, 1.0e6 + 0.4e6 * np.tanh(rs - 1.0))
    b = np.where(inside, 0.0, 2.0 / rs ** 3)
    omega = 8.93e3 * np.sqrt(ne) * 2.0 * np.pi
    session.set_omega_cube(omega, g, g, g)
    session.set_field_cubes(g, g, g, ne, te, b)
    c1 = -x_fov + 2.0 * x_fov * np.arange(n_pix) / (n_pix - 1)
    Xi, Yi = np.meshgrid(c1, c1)
    xs, ys = Xi.ravel(), Yi.ravel()
    zs = np.sqrt(np.abs(4.0 * z_obs * z_obs - xs * xs - ys * ys)) / 2.0
    area = (2.0 * x_fov / n_pix * 6.957e10) ** 2
    tb, vi, st = session.render_map(xs, ys, zs, [(freq, 6e-3 * np.sqrt(100e6 / freq), 3000, 6)], pixel_area_cm2=area,
                                    em_flag=5, use_bvec=False)
    assert (tb[0] > 1e4).mean() > 0.3
    assert f"nominal {st['nominal_ray_steps']} " in lines[0]
    active_c = int(lines[0].split(" active ")[1].split()[0])
    assert abs(active_c - st["active_ray_steps"]) <= 1e-3 * st["active_ray_steps"]      # ulps in the model can move an exit by a step
    np.testing.assert_allclose(img[0].ravel(), tb[0], rtol=1e-6, atol=1e-6 * tb[0].max())   # libm vs numpy in the model: ulps
    np.testing.assert_allclose(img[1].ravel(), vi[0], atol=1e-8)
