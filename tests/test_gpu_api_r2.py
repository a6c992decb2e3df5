"""GPU tests of the round-2 boundary work (run on the B200 box: ``pytest -m gpu``): caller-stream and
device semantics of a context, ray_order validation, the frozen ray's last record with the cross-sections
off, the fastGRFF device-array contract, the chunked host pipelines, cube re-use, cube export and the
image gather."""
import ctypes

import numpy as np
import pytest

import cases
import grff_checks
from raytracinggrff_b200 import synthetic

pytestmark = pytest.mark.gpu


def test_context_on_torch_default_stream_is_ordered_with_torch(oracle):
    """A context created on torch's default stream (handle 0, the legacy default stream) must run ON that
    stream: torch work queued before a call is finished when the library reads its buffers and torch work
    queued after it sees the library's writes, without any explicit synchronisation."""
    import torch
    from raytracinggrff_b200 import RaySession, _lib
    dev = torch.device("cuda", torch.cuda.current_device())
    assert torch.cuda.current_stream().cuda_stream == 0
    ctx = _lib.Context(dev.index, torch.cuda.current_stream().cuda_stream)
    assert ctx.stream == 0
    ses = RaySession(context=ctx)
    c = synthetic.corona_cube(48, 3.0)
    ses.set_omega_cube(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"])
    ses.set_field_cubes(c["x_grid"], c["y_grid"], c["z_grid"], c["ne"], c["te"], c["b"])
    xs, ys, zs, kv = synthetic.ray_launch_geometry(16, 1.44, 3.0)
    fps = [(75e6, 6e-3, 3000, 10)]
    area = (2 * 1.44 / 16 * 6.957e10) ** 2
    ref_tb, ref_vi, _ = ses.render_map(xs, ys, zs, fps, pixel_area_cm2=area)
    out = torch.empty((2, 1, 256), dtype=torch.float64, device=dev)
    big = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    for _ in range(3):
        # a long torch kernel chain that ends by poisoning `out`; the library call that follows must come after it
        for _ in range(20):
            big.normal_()
        out.fill_(float("nan"))
        ses.render_map(xs, ys, zs, fps, pixel_area_cm2=area, out_device_ptrs=(out[0].data_ptr(), out[1].data_ptr()))
        doubled = out * 2.0          # torch kernel queued right behind the library's
        host = doubled.cpu().numpy()
        assert np.array_equal(host[0], 2.0 * ref_tb) and np.array_equal(host[1], 2.0 * ref_vi)
    ses.close()


def test_module_level_calls_keep_the_callers_device():
    import torch
    from raytracinggrff_b200 import _lib, sample_model_with_rays
    if torch.cuda.device_count() < 2:
        # single GPU: the guard still has to leave the current device alone
        before = _lib.current_device()
        sample_model_with_rays("cuda", *cases.sampler_fixture(2), r_sun_cm=1.0)
        assert _lib.current_device() == before
        return
    torch.cuda.set_device(1)
    try:
        sample_model_with_rays("cuda", *cases.sampler_fixture(2), r_sun_cm=1.0)
        assert _lib.current_device() == 1 and torch.cuda.current_device() == 1
        ctx0 = _lib.Context(0)
        assert _lib.current_device() == 1          # creating / using a context of another GPU does not move the thread
        ctx0.synchronize()
        assert _lib.current_device() == 1
        ctx0.close()
        assert _lib.default_context().device == 1
    finally:
        torch.cuda.set_device(0)


def test_render_map_rejects_a_bad_ray_order(session):
    c = synthetic.corona_cube(32, 3.0)
    session.set_omega_cube(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"])
    session.set_field_cubes(c["x_grid"], c["y_grid"], c["z_grid"], c["ne"], c["te"], c["b"])
    xs, ys, zs, kv = synthetic.ray_launch_geometry(4, 1.0, 3.0)
    fps = [(75e6, 6e-3, 50, 10)]
    good = np.arange(16, dtype=np.int32)[::-1].copy()
    tb, _, _ = session.render_map(xs, ys, zs, fps, ray_order=good)
    tb0, _, _ = session.render_map(xs, ys, zs, fps)
    assert np.array_equal(tb, tb0)
    for bad in (np.full(16, 3, dtype=np.int32), np.arange(16, dtype=np.int32) + 1, np.arange(16, dtype=np.int32) - 1):
        with pytest.raises(ValueError, match="permutation"):
            session.render_map(xs, ys, zs, fps, ray_order=bad)


def test_fused_equals_staged_with_cross_sections_off(oracle, session):
    """trace_cs = 0: S = 1, so the record a frozen ray repeats after it stopped counts once (ds > 0) in
    trace -> sample -> emission.  The fused kernel must hand that record over too, whatever the other lanes
    of the warp are doing (the warp used to leave as soon as no lane was alive).  The cube has a
    region of NaN omega_pe in the rays' way: rays freeze INSIDE the cube there, where the frozen record
    samples real plasma."""
    c = synthetic.corona_cube(48, 3.0)
    w = c["omega_pe"].copy()
    g = c["x_grid"]
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    w[(np.abs(X - 0.4) < 0.5) & (np.abs(Y) < 0.6) & (np.abs(Z - 1.6) < 0.2)] = np.nan     # a slab some rays run into
    session.set_omega_cube(w, g, g, g)
    session.set_field_cubes(g, g, g, c["ne"], c["te"], c["b"])
    n = 12
    xs, ys, zs, kv = synthetic.ray_launch_geometry(n, 1.3, 3.0)
    area = (2 * 1.3 / n * 6.957e10) ** 2
    freq, dt, n_steps, stride = 120e6, 5e-3, 1200, 7
    r, s, _ = session.trace(freq, xs, ys, zs, kv, dt, n_steps, stride, False)
    frozen_inside = np.all(r[-1] == r[-2], axis=1) & np.all(np.abs(r[-1]) < 2.9, axis=1)
    assert 4 <= frozen_inside.sum() < n * n - 4, frozen_inside.sum()       # a mix of both kinds inside warps
    session.sample_traced(np.column_stack([xs, ys, zs]), 6.957e10, fetch=False)
    tb_s, vi_s = session.emission_traced(area, freq)
    tb_f, vi_f, _ = session.render_map(xs, ys, zs, [(freq, dt, n_steps, stride)], kvec_in_norm=kv,
                                       trace_crosssections=False, pixel_area_cm2=area)
    assert (tb_s[frozen_inside, 0] > 0).all()
    np.testing.assert_allclose(tb_f[0], tb_s[:, 0], rtol=1e-4)
    np.testing.assert_allclose(vi_f[0], vi_s[:, 0], atol=1e-4)
    # and the oracle chain agrees on the frozen pixels (those whose freezing point agrees: the step into the NaN
    # region is as discontinuous as the step out of the cube, see test_gpu_parity._cmp_paths)
    r_ref, _ = oracle.ray_trace(w, g, g, g, freq, xs, ys, zs, kv, dt, n_steps, stride, False)
    frozen_inside &= np.abs(r[-1] - r_ref[-1]).max(axis=1) < 1e-5
    assert frozen_inside.sum() >= 3
    smp = oracle.sample_model_with_rays_cpu(g, g, g, c["ne"], c["te"], c["b"], r_ref, np.ones(r_ref.shape[:2]),
                                            np.column_stack([xs, ys, zs]), 6.957e10)
    tb_ref, _, _ = oracle.emission_from_samples(smp, n, 1.3, freq)
    np.testing.assert_allclose(tb_f[0][frozen_inside], tb_ref.ravel()[frozen_inside], rtol=1e-4)


def test_get_mw_slice_device_arrays_fastgrff_contract(oracle):
    """The reference's call (script/resample_with_ray_tracing.py:428-446): CuPy Lparms_M / Rparms_M / Parms_M /
    RL_M, RL_M written in place ON THE DEVICE, a status array back.  CuPy is not installed here; torch CUDA
    tensors expose the same ``__cuda_array_interface__``."""
    import torch
    from raytracinggrff_b200 import get_mw_slice
    rng = np.random.default_rng(5)
    npix, nz, nf = 130, 61, 3
    P = grff_checks.random_los_batch(rng, npix, nz, theta90=False, flag=4, with_b=True)
    R = np.zeros((3, npix), order="F")
    R[0], R[1], R[2] = 2.5e17, 3e8, 0.2
    L = np.array([npix, nz, nf, 1, 0, 0], dtype=np.int32)
    RL_ref = np.zeros((7, nf, npix), order="F")
    oracle.get_mw_slice(L, R, P, None, None, None, RL_ref)

    def f_order(a):      # a torch tensor with the memory layout of a Fortran-ordered numpy array
        return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().permute(*reversed(range(a.ndim)))

    L_d = torch.from_numpy(L).cuda()
    R_d, P_d = f_order(R), f_order(P)
    RL_d = f_order(np.zeros((7, nf, npix)))
    dummy = torch.zeros((), dtype=torch.float64, device="cuda")
    assert RL_d.shape == (7, nf, npix) and not RL_d.is_contiguous()
    status = get_mw_slice(L_d, R_d, P_d, dummy, dummy, dummy, RL_d, tile_pixels=256, heap_bytes=2 << 30)
    assert isinstance(status, np.ndarray) and status.shape == (npix,) and not np.any(status != 0)
    RL = RL_d.cpu().numpy()
    scale = np.abs(RL_ref[1:]).max(axis=0, keepdims=True) + 1e-300
    assert (np.abs(RL[1:] - RL_ref[1:]) / scale).max() <= 1e-4
    np.testing.assert_allclose(RL[0], RL_ref[0], rtol=1e-14)
    # host Parms next to a device RL_M is staged; a C-ordered device RL_M is refused
    RL_d2 = f_order(np.zeros((7, nf, npix)))
    get_mw_slice(L, R, P, 0, 0, 0, RL_d2)
    assert np.array_equal(RL_d2.cpu().numpy(), RL)
    with pytest.raises(ValueError, match="Fortran"):
        get_mw_slice(L_d, R_d, P_d, dummy, dummy, dummy, torch.zeros((7, nf, npix), dtype=torch.float64, device="cuda"))
    with pytest.raises(TypeError):
        get_mw_slice(L_d, R_d, P_d, dummy, dummy, dummy, [0.0])


def test_chunked_host_pipelines_match_the_oracle(oracle, session):
    """rtgrff_sample / rtgrff_get_mw_slice move large host arrays in chunks through pinned bounce buffers
    while earlier chunks compute; the results are those of the oracle (bit for bit for the sampler) —
    including ds, which looks back across chunk boundaries for the previous valid record."""
    args = list(cases.los_sampler_case(160, 200, 64, seed=3))       # 25 600 rays x 200 records = 169 MB moved
    s_arr = args[7].copy()
    s_arr[::3, ::5] = 0.0
    s_arr[40:75, 1::2] = np.nan          # long invalid runs: ds reaches back over many records
    args[7] = s_arr
    xg, yg, zg, ne, te, b, r_record, s_arr, origin = args
    session.set_field_cubes(xg, yg, zg, ne, te, b)
    out = session.sample(r_record, s_arr, origin, 6.957e10)
    ref = oracle.sample_model_with_rays_cpu(*args, r_sun_cm=6.957e10)
    for k in ("ne", "te", "b", "valid_mask"):
        assert np.array_equal(out[k], ref[k], equal_nan=True), k
    np.testing.assert_allclose(out["ds"], ref["ds"], rtol=1e-6, atol=0)
    # and the pipeline is deterministic
    out2 = session.sample(r_record, s_arr, origin, 6.957e10)
    for k in ("ne", "te", "b", "ds", "valid_mask"):
        assert np.array_equal(out[k], out2[k], equal_nan=True), k
    # GRFF slice: 20 MB of Parms -> chunked over pixels
    rng = np.random.default_rng(2)
    npix, nz, nf = 4000, 120, 2          # 58 MB of Parms: two pixel chunks
    P = grff_checks.random_los_batch(rng, npix, nz, theta90=True, flag=5, with_b=True)
    Lp = np.array([npix, nz, nf, 1, 0, 0], dtype=np.int32)
    R = np.zeros((3, npix), order="F")
    R[0], R[1], R[2] = 2.5e17, 2e8, 0.25
    RL_ref = np.zeros((7, nf, npix), order="F")
    oracle.get_mw_slice(Lp, R, P, None, None, None, RL_ref)
    RL = np.zeros((7, nf, npix), order="F")
    status = session.get_mw_slice(Lp, R, P, RL)
    assert not status.any()
    scale = np.abs(RL_ref[1:]).max(axis=0, keepdims=True) + 1e-300
    assert (np.abs(RL[1:] - RL_ref[1:]) / scale).max() <= 1e-4


def test_drop_in_calls_reuse_the_uploaded_cube(oracle):
    """trace_ray / sample_model_with_rays called again with the same arrays skip the upload (the reference
    re-uploads per call); a changed array is uploaded again."""
    from raytracinggrff_b200 import _lib, trace_ray
    c = synthetic.corona_cube(64, 3.0)
    xs, ys, zs, kv = synthetic.ray_launch_geometry(8, 1.2, 3.0)
    kw = dict(x_grid=c["x_grid"], y_grid=c["y_grid"], z_grid=c["z_grid"], freq_hz=75e6, x_start=xs, y_start=ys, z_start=zs,
              kvec_in_norm=kv, dt=6e-3, n_steps=400, record_stride=20)
    ctx = _lib.default_context()
    r1, _ = trace_ray("cuda", c["omega_pe"], **kw)
    n1 = ctx.launch_count
    r2, _ = trace_ray("cuda", c["omega_pe"], **kw)
    n2 = ctx.launch_count
    assert np.array_equal(r1, r2)
    assert n2 - n1 < n1 and n2 - n1 <= 2            # trace + layout kernel only: no cube build
    w2 = c["omega_pe"] * 1.3
    r3, _ = trace_ray("cuda", w2, **kw)
    assert ctx.launch_count - n2 > n2 - n1          # cube rebuilt
    r_ref, _ = oracle.ray_trace(w2, c["x_grid"], c["y_grid"], c["z_grid"], 75e6, xs, ys, zs, kv, 6e-3, 400, 20)
    assert np.nanmax(np.abs(r3 - r_ref)) < 1e-5 and np.nanmax(np.abs(r3 - r1)) > 1e-4


def test_export_cubes_round_trip(session):
    c = synthetic.corona_cube(40, 3.0, active_region=True)
    g3 = (c["x_grid"], c["y_grid"], c["z_grid"])
    session.set_field_cubes(*g3, c["ne"], c["te"], c["b"], c["bx"], c["by"], c["bz"])
    session.set_omega_cube(c["omega_pe"], *g3)
    e = session.export_cubes(omega_pe=True, fields=True, bvec=True)
    assert np.array_equal(e["omega_pe"], c["omega_pe"])
    for k in ("ne", "te", "b", "bx", "by", "bz"):
        assert np.array_equal(e[k], c[k].astype(np.float32)), k


def test_gather_image_single_rank_and_shard_layout(session):
    """World size 1: the gather is the row placement alone.  The slab of a rank lists its rows in
    rtgrff_shard_rows order; emulate 3 ranks' slabs on one GPU by checking the placement kernel through the
    single-rank path on each rank's rows (the NCCL exchange itself is exercised by bench.py --gpus N)."""
    import torch
    from raytracinggrff_b200 import dist as rdist
    n_planes, n_rows, n_cols = 4, 37, 20
    session.comm_init(1, 0)
    img = np.arange(n_planes * n_rows * n_cols, dtype=np.float64).reshape(n_planes, n_rows, n_cols)
    slab = torch.from_numpy(img).cuda()
    out = np.zeros_like(img)
    session.gather_image(slab.data_ptr(), n_planes, n_rows, n_cols, root=0, out=out)
    assert np.array_equal(out, img)
    pinned = torch.empty(img.shape, dtype=torch.float64, pin_memory=True)
    session.gather_image(slab.data_ptr(), n_planes, n_rows, n_cols, root=0, out=pinned.numpy())
    assert np.array_equal(pinned.numpy(), img)
    dev_out = torch.zeros_like(slab)
    session.gather_image(slab.data_ptr(), n_planes, n_rows, n_cols, root=0, out_device_ptr=dev_out.data_ptr())
    assert np.array_equal(dev_out.cpu().numpy(), img)
    for world in (2, 3, 8):
        for n in (37, 200, 1024):
            rows_all = np.concatenate([rdist.c_shard_rows(n, world, r)[0] for r in range(world)])
            assert np.array_equal(np.sort(rows_all), np.arange(n))
            for r in range(world):
                rows, mr = rdist.c_shard_rows(n, world, r)
                assert np.array_equal(rows, rdist.rows_of_rank(n, world, r)) and mr == rdist.max_rows_per_rank(n, world)


@pytest.mark.parametrize("backend", ["fused", "device", "fastgrff"])
def test_workers_split_rays_over_gpus_without_changing_the_map(backend):
    """``--workers N`` (script/resample_with_ray_tracing.py:333-352: contiguous ray chunks, concatenated): here the
    chunks go to the visible GPUs, each with its own context and cube replica.  Rays are independent, so the map
    is the single-worker map bit for bit — also when the chunks do not divide the image evenly."""
    from raytracinggrff_b200.workflow import run_ray_tracing_emission
    c = synthetic.corona_cube(48, 3.0)
    kw = dict(N_pix=10, X_fov=1.3, freq_hz=90e6, z_observer=3.0, dt=6e-3, n_steps=1500, record_stride=8,
              grff_backend=backend, verbose=False)
    one = run_ray_tracing_emission(c, n_workers=1, **kw)
    many = run_ray_tracing_emission(c, n_workers=3, **kw)
    assert (one["emission_cube"] > 0).mean() > 0.3
    assert np.array_equal(one["emission_cube"], many["emission_cube"])
    assert np.array_equal(one["emission_polVI_cube"], many["emission_polVI_cube"])


def test_fp32_voxel_path_against_forced_fp64(session):
    """The float32 voxel evaluation of the per-ray kernels (csrc/grff_fast.cuh) against the same kernels with every
    voxel forced through FP64, on maps where the cut-offs matter: low frequencies (rays turn where v -> 1, the
    voxels around the turning point take the FP64 fallback) and GHz frequencies over the active region (strong B:
    X-mode cut-off, gyro-resonance guard).  Both the fused and the staged (emission_traced) consumers."""
    c = synthetic.corona_cube(96, 3.0, active_region=True)
    g3 = (c["x_grid"], c["y_grid"], c["z_grid"])
    session.set_omega_cube(c["omega_pe"], *g3)
    session.set_field_cubes(*g3, c["ne"], c["te"], c["b"], c["bx"], c["by"], c["bz"])
    n = 24
    xs, ys, zs, kv = synthetic.ray_launch_geometry(n, 1.3, 3.0)
    area = (2 * 1.3 / n * 6.957e10) ** 2
    fps = [dict(freq_hz=f, **synthetic.frequency_scaled_params(f)) for f in (40e6, 150e6, 1.2e9)]
    out = {}
    try:
        for mode in (0, 1):
            session.ctx.set_grff64(mode)
            tb, vi, _ = session.render_map(xs, ys, zs, fps, pixel_area_cm2=area, em_flag=4, use_bvec=True)
            p = fps[1]
            session.trace(p["freq_hz"], xs, ys, zs, kv, p["dt"], p["n_steps"], p["record_stride"], True, 2.0, fetch=False)
            session.sample_traced(np.column_stack([xs, ys, zs]), 6.957e10, fetch=False)
            tb_s, vi_s = session.emission_traced(area, p["freq_hz"])
            out[mode] = (tb, vi, tb_s, vi_s)
    finally:
        session.ctx.set_grff64(0)
    for a, b in ((out[0][0], out[1][0]), (out[0][2], out[1][2])):
        on = b != 0
        assert on.mean() > 0.3 and np.array_equal(a != 0, on)
        rel = np.abs(a[on] - b[on]) / b[on]
        assert rel.max() < 2e-5, rel.max()
    assert np.abs(out[0][1] - out[1][1]).max() < 2e-5 and np.abs(out[0][3] - out[1][3]).max() < 2e-5
    assert not np.array_equal(out[0][0], out[1][0])          # the two paths are different arithmetic


@pytest.mark.parametrize("world,n_rows", [(2, 37), (3, 200), (8, 1000), (4, 64), (8, 2048)])
def test_place_rows_with_ragged_shares(session, world, n_rows):
    """The root-side reorder of the image gather for shares of unequal size (the benchmarks only ever have equal
    ones): every rank's slab, padded to the largest share, in; the image with every row at its place out."""
    import torch
    from raytracinggrff_b200 import _lib, dist as rdist
    n_planes, n_cols = 3, 11
    rng = np.random.default_rng(world * 1000 + n_rows)
    img = rng.random((n_planes, n_rows, n_cols))
    mr = rdist.max_rows_per_rank(n_rows, world)
    gathered = np.full((world, n_planes, mr, n_cols), np.nan)
    for r in range(world):
        rows, mr_c = rdist.c_shard_rows(n_rows, world, r)
        assert mr_c == mr
        gathered[r, :, :len(rows)] = img[:, rows]
    g_d = torch.from_numpy(gathered).cuda()
    out_d = torch.zeros((n_planes, n_rows, n_cols), dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().rtgrff_place_rows(session.ctx.handle, ctypes.c_void_p(g_d.data_ptr()), world, n_planes, n_rows,
                                             n_cols, ctypes.c_void_p(out_d.data_ptr())))
    assert np.array_equal(out_d.cpu().numpy(), img)


def test_stepper_without_the_polynomial_cube_is_bit_identical(session, monkeypatch):
    """Cubes too large for the 8x cell-major polynomial copy are stepped from the node cube, the cell's polynomial
    differenced on the fly (load_cell_nodes): same arithmetic, same bits — paths and maps."""
    c = synthetic.corona_cube(64, 3.0, active_region=True)
    g3 = (c["x_grid"], c["y_grid"], c["z_grid"])
    xs, ys, zs, kv = synthetic.ray_launch_geometry(12, 1.3, 3.0)
    area = (2 * 1.3 / 12 * 6.957e10) ** 2
    fps = [(90e6, 6e-3, 2500, 7), (600e6, 2.4e-3, 4000, 2)]
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("RTGRFF_POLY_CUBE", flag)
        session.set_omega_cube(c["omega_pe"], *g3)
        session.set_field_cubes(*g3, c["ne"], c["te"], c["b"], c["bx"], c["by"], c["bz"])
        r, s, act = session.trace(90e6, xs, ys, zs, kv, 6e-3, 2500, 7, True, 2.0)
        tb, vi, st = session.render_map(xs, ys, zs, fps, pixel_area_cm2=area, em_flag=4, use_bvec=True)
        out[flag] = (r, s, act, tb, vi, st)
    monkeypatch.delenv("RTGRFF_POLY_CUBE")
    session.set_omega_cube(c["omega_pe"], *g3)            # leave the shared session with the default layout
    for a, b in zip(out["1"], out["0"]):
        if isinstance(a, np.ndarray):
            assert np.array_equal(a, b, equal_nan=True)
        else:
            assert a == b
    assert (out["1"][3] > 0).mean() > 0.3
