"""CPU tier: host-side logic (PyTorch ingest, C-ABI surface, slab planning and query routing).
No GPU compute is called here."""
import ctypes
import os
import re
import socket
import warnings

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden

from arbinterp_b200 import _lib
from arbinterp_b200.ingest import FieldError, ingest_field, norm_plane, sorted_field
from arbinterp_b200.sharding import exchange_and_query, owner_ranks, plan_slabs, push_sharded, slab_planes


# ----------------------------------------------------------------------------------------- ingest
@pytest.mark.parametrize("name,d", [("tri_12x10x9", 3), ("quad_8x7x7x6", 4), ("tri_scalar_9x8x11", 3)])
def test_ingest_matches_reference_geometry(name, d):
    g = load_golden(name)
    planes, geo = ingest_field(g["field"], d)             # rows are shuffled in the fixture
    assert geo.h == list(g["h"]) and geo.int_min == list(g["int_min"]) and geo.int_max == list(g["int_max"])
    assert geo.ncell == list(g["ncell_axis"]) and geo.nc == int(g["nc"])
    srt = g["sorted_field"]                                # the reference's sorted inputfield
    assert np.array_equal(sorted_field(planes, geo).numpy(), srt)
    for c in range(planes.shape[0]):
        assert np.array_equal(planes[c].reshape(-1).numpy(), srt[:, d + c])


def test_ingest_does_not_modify_input_and_accepts_tensors():
    g = load_golden("tri_12x10x9")
    f = g["field"].copy()
    p1, _ = ingest_field(f, 3)
    assert np.array_equal(f, g["field"])
    p2, _ = ingest_field(torch.from_numpy(f), 3)
    assert torch.equal(p1, p2)


def test_norm_plane_matches_numpy():
    """Same operation order as np.linalg.norm(axis=1) (A.py:58).  torch's CPU sqrt is a vectorised
    approximation that is 1 ulp off in places, so this CPU check allows 1 ulp; the CUDA sqrt is
    IEEE and the gpu-marked twin of this test (test_gpu_parity.py) asserts bit equality."""
    g = load_golden("tri_12x10x9")
    planes, _ = ingest_field(g["field"], 3)
    ref = np.linalg.norm(g["sorted_field"][:, 3:], axis=1)
    got = norm_plane(planes).reshape(-1).numpy()
    assert np.max(np.abs(got - ref) / np.spacing(ref)) <= 1.0


def test_ingest_rejects_bad_grids():
    g = load_golden("tri_12x10x9")
    f = g["field"]
    with pytest.raises(FieldError, match="full grid"):
        ingest_field(f[:-1], 3)                            # a missing point
    dup = f.copy(); dup[0, :3] = dup[1, :3]
    with pytest.raises(FieldError, match="full grid"):
        ingest_field(dup, 3)
    x = np.linspace(0, 1, 3)
    Z, Y, X = np.meshgrid(x, x, x, indexing="ij")
    with pytest.raises(FieldError, match="at least 4"):
        ingest_field(np.stack([X.ravel(), Y.ravel(), Z.ravel(), X.ravel()], 1), 3)
    bad = f.copy(); bad[3, 1] = np.nan
    with pytest.raises(FieldError, match="NaN"):
        ingest_field(bad, 3)


def test_ingest_warns_on_irregular_spacing():
    x = np.array([0.0, 1.0, 2.0, 3.5, 4.0]); y = np.arange(4.0)
    Z, Y, X = np.meshgrid(y, y, x, indexing="ij")
    with pytest.warns(RuntimeWarning, match="not evenly spaced"):
        ingest_field(np.stack([X.ravel(), Y.ravel(), Z.ravel(), X.ravel()], 1), 3)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        ingest_field(load_golden("tri_12x10x9")["field"], 3)   # linspace grids pass


def test_csv_loader_matches_genfromtxt(tmp_path):
    """arbinterp_b200.io.load_field_csv == np.genfromtxt(delimiter=',') on the reference's file format."""
    from arbinterp_b200.io import load_field_csv, save_field_csv
    g = load_golden("tri_12x10x9")
    path = str(tmp_path / "field.csv")
    save_field_csv(path, g["field"])
    a = load_field_csv(path)
    assert a.dtype == np.float64 and a.flags.c_contiguous
    assert np.array_equal(a, np.genfromtxt(path, delimiter=",")) and np.array_equal(a, g["field"])   # %.18e round-trips


# ----------------------------------------------------------------------------------------- C-ABI
def _declared_functions():
    text = open(os.path.join(ROOT, "include", "arbinterp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(arb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_functions()
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert b"arbinterp_b200" in lib.arb_version()


def test_exact_matrices_against_reference_fixtures():
    from oracle.arb_oracle import a_matrix, difference_matrix, hermite_matrix
    m = load_golden("matrices")
    for d, key, sc in ((3, "A3_times8", 8), (4, "A4_times16", 16)):
        assert np.array_equal(_lib.get_matrix(d, "A") * sc, m[key].astype(np.float64))
        B = m["B3" if d == 3 else "B4"].astype(np.float64)
        assert np.array_equal(_lib.get_matrix(d, "invB") @ B, np.eye(4 ** d))     # exact integer inverse
        assert np.array_equal(_lib.get_matrix(d, "D"), difference_matrix(d))
        assert np.array_equal(_lib.get_matrix(d, "A", False), a_matrix(d, False))
        assert np.array_equal(hermite_matrix(d), B)


def test_separable_build_phases_on_host(tmp_path):
    """The phases of the separable build kernels (arb_build_sep.cuh, __host__ __device__) run thread by thread
    on the CPU and reproduce A f for every cell, in 3-D and in 4-D with and without the A.py:860 quirk."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "sep_host_emul")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(ROOT, "tests", "host_emul", "sep_host_emul.cu"),
                    os.path.join(ROOT, "arbinterp_b200", "csrc", "arb_core.cu")], check=True, timeout=600)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    errs = [float(x) for x in re.findall(r"max scaled error ([0-9.eE+-]+)", out.stdout)]
    assert out.returncode == 0, out.stdout
    assert len(errs) >= 10 and max(errs) <= 1e-12


def test_table_free_quadcubic_math_on_host(tmp_path):
    """The __host__ __device__ pieces of the 4-D table-free kernel (plane contraction with rotated row order,
    parity sums, rank-16 corner term), combined as the kernel's four lanes combine them, equal the monomial
    evaluation of A f with and without the A.py:860 quirk."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "gridfree4_host_emul")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(ROOT, "tests", "host_emul", "gridfree4_host_emul.cu"),
                    os.path.join(ROOT, "arbinterp_b200", "csrc", "arb_core.cu")], check=True, timeout=600)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    errs = [float(x) for x in re.findall(r"max scaled error ([0-9.eE+-]+)", out.stdout)]
    assert out.returncode == 0, out.stdout
    assert len(errs) >= 4 and max(errs) <= 1e-12


def test_interleaved_grid_4d_math_on_host(tmp_path):
    """The __host__ __device__ pieces of the 4-D interleaved table-free kernel (arb_gridil4.cuh: one lane per z-plane,
    one pass per t-plane, parity sums for the A.py:860 term and the two lane exchanges) equal the monomial evaluation of
    alpha = A f for every component, 'vector' and 'both', with and without the quirk."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "gridil4_host_emul")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "--expt-relaxed-constexpr", "-o", exe,
                    os.path.join(ROOT, "tests", "host_emul", "gridil4_host_emul.cu"),
                    os.path.join(ROOT, "arbinterp_b200", "csrc", "arb_core.cu")], check=True, timeout=600)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    errs = [float(x) for x in re.findall(r"max scaled error ([0-9.eE+-]+)", out.stdout)]
    assert out.returncode == 0, out.stdout
    assert len(errs) >= 5 and max(errs) <= 1e-12


def test_node_table_math_on_host(tmp_path):
    """The __host__ __device__ pieces of the node (Hermite) table (arb_nodes.cuh: the central-difference stencil of the
    build, the 3-D evaluation, the four-lane 4-D evaluation with the A.py:860 term) equal the monomial evaluation of
    alpha = A f, with and without the quirk."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "nodes_host_emul")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "--expt-relaxed-constexpr", "-o", exe,
                    os.path.join(ROOT, "tests", "host_emul", "nodes_host_emul.cu"),
                    os.path.join(ROOT, "arbinterp_b200", "csrc", "arb_core.cu")], check=True, timeout=600)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    errs = [float(x) for x in re.findall(r"max scaled error ([0-9.eE+-]+)", out.stdout)]
    assert out.returncode == 0, out.stdout
    assert len(errs) >= 7 and max(errs) <= 1e-12


def _compile_c_client(out):
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    cuda = "/usr/local/cuda"
    if not gcc or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA runtime headers are not available")
    libdir = os.path.join(ROOT, "arbinterp_b200")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-I",
                    os.path.join(cuda, "include"), "-o", out, os.path.join(ROOT, "tests", "c_abi", "c_smoke.c"),
                    "-L", libdir, "-larbinterp_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm",
                    "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda, "lib64")], check=True, timeout=300)
    return out


def test_header_is_plain_c_and_c_client_links(tmp_path):
    """The drop-in boundary is a C ABI: the header compiles as strict C99 and a plain-C client
    (tests/c_abi/c_smoke.c) builds and links against the shared library (it runs in the GPU suite)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                    os.path.join(ROOT, "include", "arbinterp_b200.h")], check=True, timeout=120)
    _lib.load()                                   # makes sure the library is built
    _compile_c_client(str(tmp_path / "c_smoke"))


def test_argument_validation_without_gpu():
    lib = _lib.load()
    g = _lib.ArbGeom()
    g.d, g.ncomp = 3, 1
    for a in range(3):
        g.ncell[a], g.h[a], g.int_min[a], g.int_max[a] = 5, 0.1, 0.0, 0.5
    g.slab_lo, g.slab_hi = 0, 5
    assert lib.arb_query(ctypes.byref(g), None, _lib.MODE_NORM, None, 0, 3, None, None, None, None, None, None, None) == 0
    assert lib.arb_query(ctypes.byref(g), None, _lib.MODE_BOTH, None, 4, 3, None, None, None, None, None, None, None) != 0
    assert b"components" in lib.arb_last_error()
    assert lib.arb_query(ctypes.byref(g), None, _lib.MODE_NORM, None, 4, 2, None, None, None, None, None, None, None) != 0
    assert b"ldq" in lib.arb_last_error()
    n = (ctypes.c_int64 * 4)(8, 8, 8, 1)
    assert lib.arb_build_coeffs(5, None, 1, ctypes.byref(n), None, 1, None) != 0
    assert lib.arb_get_matrix(2, 0, 1, None) != 0
    old = lib.arb_set_query_variant(1)
    assert lib.arb_set_query_variant(old) == 1
    # the fused legs of slab-sharded queries: argument checks come before any CUDA call
    hi = (ctypes.c_int64 * 2)(2, 5)
    vp2 = (ctypes.c_void_p * 2)(None, None)
    assert lib.arb_route_rows(ctypes.byref(g), None, 10, 3, hi, 2, 0, vp2, vp2, 4, None, None, None, None) != 0      # n > seg_cap
    assert b"arb_route_rows" in lib.arb_last_error()
    assert lib.arb_route_rows(ctypes.byref(g), None, 0, 3, hi, 2, 5, vp2, vp2, 4, None, None, None, None) != 0       # rank out of range
    assert lib.arb_query_inbox(ctypes.byref(g), None, _lib.MODE_NORM, None, None, 4, vp2, 2, 6, None) != 0
    assert b"arb_query_inbox" in lib.arb_last_error()
    assert lib.arb_query_gridil(ctypes.byref(g), None, _lib.MODE_NORM, None, 4, 3, None, None, None, None, None, None, None) != 0


def test_wrong_shape_exits_like_reference():
    from arbinterp_b200 import quadcubic, tricubic
    with pytest.raises(SystemExit, match="N x 4 or N x 6"):
        tricubic(np.zeros((10, 5)), "quiet")                   # A.py:104
    with pytest.raises(SystemExit, match="N x 5 or N x 7"):
        quadcubic(np.zeros((10, 4)), "quiet")                  # A.py:723


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from arbinterp_b200 import tricubic
    with pytest.raises(_lib.ArbError, match="no CPU fallback"):
        tricubic(load_golden("tri_12x10x9")["field"], "quiet")


def test_drop_in_import_paths():
    import ARBInterp.ARBInterp as a
    import ARBTools.ARBInterp as b
    import arbinterp_b200
    assert a.tricubic is arbinterp_b200.tricubic and b.quadcubic is arbinterp_b200.quadcubic
    assert a.__version__ == "1.8"


# ----------------------------------------------------------------------------------------- sharding
def test_plan_slabs():
    assert plan_slabs(61, 8) == [(0, 8), (8, 16), (16, 24), (24, 32), (32, 40), (40, 47), (47, 54), (54, 61)]
    assert plan_slabs(3, 1) == [(0, 3)]
    assert plan_slabs(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert slab_planes((8, 16)) == (8, 19)                  # 1 halo plane below the first corner, 2 above
    with pytest.raises(ValueError):
        plan_slabs(0, 2)


def test_owner_ranks_follow_cell_location():
    slabs = plan_slabs(10, 3)                              # (0,4) (4,7) (7,10)
    t0, h = 0.5, 0.25
    t1 = t0 + 10 * h
    t = torch.tensor([t0, t0 + 3.999 * h, t0 + 4 * h, t0 + 6.5 * h, t0 + 7 * h, t1, t0 - 1e-9, t1 + 1.0, float("nan")],
                     dtype=torch.float64)
    own = owner_ranks(t, t0, t1, h, slabs)
    assert own.tolist() == [0, 0, 1, 1, 2, 2, 0, 0, 0]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _routing_worker(rank, world, port, field, q_all, out_dir):
    import torch.distributed as dist
    from oracle.arb_oracle import OracleInterp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = 4
    ora = OracleInterp(field, d, mode="both")
    geo = ora.geo
    slabs = plan_slabs(geo.ncell_axis[d - 1], world)
    lo, hi = slabs[rank]
    seen = []

    def evaluate(rows):
        # the local "slab": must only ever be asked about its own layers (or unowned NaN rows on rank 0)
        r = rows.numpy().copy()
        layer = np.floor((r[:, d - 1] - geo.int_min[d - 1]) / geo.h[d - 1])
        inside = (r[:, d - 1] >= geo.int_min[d - 1]) & (r[:, d - 1] <= geo.int_max[d - 1])
        seen.append(bool(np.all(~inside | ((layer >= lo) & (layer < hi)) | (rank == world - 1))))
        comps, norms, grads = ora.query(r, exact_gemv=True)
        return torch.from_numpy(np.hstack([comps, norms, grads]))

    mine = torch.from_numpy(q_all[rank::world].copy())
    owner = owner_ranks(mine[:, d - 1], geo.int_min[d - 1], geo.int_max[d - 1], geo.h[d - 1], slabs)
    out = exchange_and_query(mine, owner, evaluate, 3 + 1 + d)
    np.save(os.path.join(out_dir, f"out{rank}.npy"), out.numpy())
    # the addressing of the fused return leg (arb_query_routed): rows arrive in segments by sender together with the row
    # number they have in the sender's batch; "storing" every result at (sender, that row) -- emulated here with an
    # all-gather -- must reproduce the caller's order without any re-ordering pass
    from arbinterp_b200.sharding import route_rows
    recv, _, _, recv_split, home_rows = route_rows(mine, owner, with_home_rows=True)
    res = evaluate(recv)
    home = torch.repeat_interleave(torch.arange(world), torch.tensor(recv_split))
    boxes = [None] * world
    dist.all_gather_object(boxes, (home.numpy(), home_rows.numpy(), res.numpy()))
    landed = np.full(out.shape, np.inf)
    for h, rows, vals in boxes:
        sel = h == rank
        landed[rows[sel]] = vals[sel]
    seen.append(bool(np.array_equal(landed, out.numpy(), equal_nan=True)))
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array(seen))
    dist.barrier()
    dist.destroy_process_group()


def _bcast_worker(rank, world, port, field, out_dir):
    import torch.distributed as dist
    from arbinterp_b200.sharding import broadcast_ingested
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    got = broadcast_ingested(field if rank == 0 else None, 4, src=0)
    g = got.geo
    np.savez(os.path.join(out_dir, f"ing{rank}.npz"), planes=got.planes.numpy(), h=g.h, lo=g.int_min, hi=g.int_max, npts=g.npts)
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_ingested_world2_gloo(tmp_path):
    """Ingest once on rank 0, broadcast axes + dense planes: every rank ends up with the planes and the
    bit-identical geometry a local ingest of the raw rows would give."""
    import torch.multiprocessing as mp
    g = load_golden("quad_8x7x7x6")
    mp.spawn(_bcast_worker, args=(2, _free_port(), g["field"], str(tmp_path)), nprocs=2, join=True)
    planes, geo = ingest_field(g["field"], 4)
    for rank in range(2):
        z = np.load(tmp_path / f"ing{rank}.npz")
        assert np.array_equal(z["planes"], planes.numpy()) and list(z["npts"]) == geo.npts
        assert list(z["h"]) == geo.h and list(z["lo"]) == geo.int_min and list(z["hi"]) == geo.int_max


def test_query_routing_world2_gloo(tmp_path):
    """World-size-2 exchange on CPU: every row reaches the rank owning its t-layer and the results
    come back in the caller's order, identical to a single-process evaluation."""
    import torch.multiprocessing as mp
    from oracle.arb_oracle import OracleInterp
    g = load_golden("quad_8x7x7x6")
    field, q_all = g["field"], g["both_q_in"][:, :4].copy()
    world = 2
    mp.spawn(_routing_worker, args=(world, _free_port(), field, q_all, str(tmp_path)), nprocs=world, join=True)
    ref = np.hstack(OracleInterp(field, 4, mode="both").query(q_all.copy(), exact_gemv=True))
    for rank in range(world):
        got = np.load(tmp_path / f"out{rank}.npy")
        assert np.array_equal(got, ref[rank::world], equal_nan=True)
        assert np.load(tmp_path / f"ok{rank}.npy").all()


# ---- sharded push: particle migration between slab owners (CPU stand-in for the resumable kernel)
_PUSH = dict(t_min=0.0, t_max=6.0, h=1.0, n_layers=6, x_lim=5.0, dt=0.37, nsteps=11)


def _fake_advance(p, v, step, lo, hi):
    """What arb_push_steps does, with a force that depends on the t layer (so a step evaluated by the wrong rank or
    twice would show): drift in x and t, parked on leaving layers [lo, hi), lost on leaving the volume."""
    c = _PUSH
    for n in range(p.shape[0]):
        while step[n] <= c["nsteps"]:
            x, t = float(p[n, 0]), float(p[n, 1])
            if not (c["t_min"] <= t <= c["t_max"]) or abs(x) > c["x_lim"] or np.isnan(x):
                p[n, 0] = float("nan"); v[n, :] = float("nan"); step[n] = c["nsteps"] + 1
                break
            layer = int(np.floor((t - c["t_min"]) / c["h"]))
            if layer >= c["n_layers"]:
                p[n, 0] = float("nan"); v[n, :] = float("nan"); step[n] = c["nsteps"] + 1   # upper edge: NaN (DESIGN)
                break
            if not (lo <= layer < hi):
                break                                                  # parked unchanged
            a = 0.1 * (layer + 1)
            if step[n] > 0:
                v[n, 0] += 0.5 * c["dt"] * a
            if step[n] < c["nsteps"]:
                v[n, 0] += 0.5 * c["dt"] * a
                p[n, 0] += c["dt"] * v[n, 0]
                p[n, 1] += c["dt"]
            step[n] += 1


def _push_worker(rank, world, port, pos_all, vel_all, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = _PUSH
    slabs = plan_slabs(c["n_layers"], world)
    lo, hi = slabs[rank]
    pos = torch.from_numpy(pos_all[rank::world].copy())
    vel = torch.from_numpy(vel_all[rank::world].copy())
    rounds = []

    def advance(p, v, step):
        rounds.append(int(p.shape[0]))
        _fake_advance(p, v, step, lo, hi)

    lost = push_sharded(pos, vel, c["nsteps"], lambda p: owner_ranks(p[:, 1], c["t_min"], c["t_max"], c["h"], slabs), advance)
    np.savez(os.path.join(out_dir, f"push{rank}.npz"), pos=pos.numpy(), vel=vel.numpy(), lost=lost, rounds=len(rounds))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_push_migration_gloo(tmp_path, world):
    """push_sharded over gloo: particles drift through the t slabs of several ranks (and some out of the volume);
    positions, velocities and loss counts equal a single-process run of the same stepping rule, rows come home in
    their original order, and the run needs more than one round (so particles really migrated)."""
    import torch.multiprocessing as mp
    rng = np.random.default_rng(3)
    n = 301
    pos = np.stack([rng.uniform(-4.5, 4.5, n), rng.uniform(0.0, 5.9, n)], axis=1)
    pos[5, 1] = 7.5                                  # starts outside the volume
    pos[6, 0] = np.nan
    vel = np.stack([rng.normal(0, 0.8, n), np.zeros(n), np.zeros(n)], axis=1)
    mp.spawn(_push_worker, args=(world, _free_port(), pos, vel, str(tmp_path)), nprocs=world, join=True)
    p_ref, v_ref = torch.from_numpy(pos.copy()), torch.from_numpy(vel.copy())
    step = torch.zeros(n, dtype=torch.int64)
    _fake_advance(p_ref, v_ref, step, 0, _PUSH["n_layers"])
    assert bool((step == _PUSH["nsteps"] + 1).all())
    total_lost, rounds = 0, []
    for rank in range(world):
        z = np.load(tmp_path / f"push{rank}.npz")
        assert np.array_equal(z["pos"], p_ref.numpy()[rank::world], equal_nan=True)
        assert np.array_equal(z["vel"], v_ref.numpy()[rank::world], equal_nan=True)
        assert int(z["lost"]) == int(np.isnan(p_ref.numpy()[rank::world, 0]).sum())
        rounds.append(int(z["rounds"]))
        total_lost += int(z["lost"])
    assert 0 < total_lost < n and max(rounds) > 1


# ---- drop-in semantics of the sharded / replicated wrappers (CPU stand-in for the CUDA class, gloo, world 2)
class _OracleBackedQuad:
    """What SlabShardedInterp / ReplicatedInterp need from ``quadcubic`` -- constructor taking the broadcast
    IngestedField (+ slab), ``Query(rows)`` on tensors, ``_last_cells``, ``_geo``, ``_mode``, ``_device``, ``nc`` --
    with the arithmetic done by the numpy oracle, so the routing, the in-place NaN rows and the global ``queryInds``
    can be checked on CPU."""
    _d = 4

    def __init__(self, field, *args, slab=None, device=None, mode="vector", **kwargs):
        from arbinterp_b200.ingest import IngestedField
        from oracle.arb_oracle import OracleInterp
        assert isinstance(field, IngestedField) and "quiet" in args
        geo = field.geo
        planes = field.planes.numpy()
        vals = dict(x=planes[0].ravel(), y=planes[1].ravel(), z=planes[2].ravel())
        vals["n"] = np.linalg.norm(np.stack([vals[k] for k in "xyz"], axis=1), axis=1)
        self._ora = OracleInterp.from_planes([a.numpy() for a in geo.axes], vals, mode, dense=True)
        self._geo, self._mode, self._device = geo, mode, torch.device("cpu")
        self._slab = slab if slab is not None else (0, geo.ncell[3])
        self.nc = geo.nc
        self.foreign_rows = 0

    def Query(self, q):
        r = (q.numpy() if isinstance(q, torch.Tensor) else q)
        r = np.ascontiguousarray(r[:, :4], dtype=np.float64).copy()
        res = self._ora.query(r, exact_gemv=True)
        inds = self._ora.query_inds
        layer = inds // (self.nc // self._geo.ncell[3])
        self.foreign_rows += int(((inds < self.nc) & ((layer < self._slab[0]) | (layer >= self._slab[1]))).sum())
        self._last_cells = torch.from_numpy(inds.astype(np.int64))
        if isinstance(q, torch.Tensor):
            return tuple(torch.from_numpy(np.ascontiguousarray(x)) for x in res)
        if isinstance(q, np.ndarray):
            q[np.isnan(r).all(axis=1) & ~np.isnan(q[:, :4]).all(axis=1)] = np.nan
        return res


def _dropin_worker(rank, world, port, field, q_all, out_dir):
    import torch.distributed as dist
    from arbinterp_b200.sharding import ReplicatedInterp, SlabShardedInterp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    src = field if rank == 0 else None
    sh = SlabShardedInterp(_OracleBackedQuad, src, "quiet", mode="both", device="cpu")
    mine = q_all[rank::world].copy()                         # numpy in, extra columns, rows of every kind
    before = mine.copy()
    comps, norms, grads = sh.Query(mine)                     # numpy out
    assert isinstance(comps, np.ndarray) and sh.local.foreign_rows == 0
    inds = sh.queryInds.copy()                               # side effect of THIS call (A.py:1088-1090)
    as_tensor = sh.Query(torch.from_numpy(before.copy()))    # tensor in -> tensors out
    assert all(isinstance(t, torch.Tensor) for t in as_tensor)
    f32 = sh.Query(torch.from_numpy(before[:, :4].astype(np.float32)))   # float32 rows must not be mis-read
    sh.Query(np.empty((0, 4)))                               # an empty share is a legal collective call
    rep = ReplicatedInterp(_OracleBackedQuad, src, "quiet", mode="both", device="cpu")
    mine2 = before.copy()
    r2 = rep.Query(mine2)
    np.savez(os.path.join(out_dir, f"dropin{rank}.npz"), out=np.hstack([comps, norms, grads]), q_after=mine,
             inds=inds, tens=np.hstack([t.numpy() for t in as_tensor]), f32=np.hstack([t.numpy() for t in f32]),
             rep=np.hstack(r2), rep_q=mine2, rep_inds=rep.local._last_cells.numpy(), nc=rep.nc)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_and_replicated_query_are_dropin_world2_gloo(tmp_path):
    """SlabShardedInterp.Query / ReplicatedInterp.Query take THIS rank's rows as numpy and behave like the reference's
    rQuery3 on them (A.py:1190-1258): same outputs as a single-process evaluation, the caller's out-of-volume rows
    NaN-overwritten in place across all columns (A.py:1069-1076), queryInds = global cell index, nc for NaN rows
    (A.py:1088-1090); rows only ever reach the rank that owns their t layer."""
    import torch.multiprocessing as mp
    from oracle.arb_oracle import OracleInterp
    g = load_golden("quad_8x7x7x6")
    field, q_all = g["field"], g["both_q_in"].copy()
    assert q_all.shape[1] > 4                                # extra columns present
    world = 2
    mp.spawn(_dropin_worker, args=(world, _free_port(), field, q_all, str(tmp_path)), nprocs=world, join=True)
    ora = OracleInterp(field, 4, mode="both")
    q_ref = q_all.copy()
    ref = np.hstack(ora.query(q_ref, exact_gemv=True))
    assert np.array_equal(q_ref, g["both_q_after"], equal_nan=True)
    for rank in range(world):
        z = np.load(tmp_path / f"dropin{rank}.npz")
        for key in ("out", "tens", "rep"):
            assert np.array_equal(z[key], ref[rank::world], equal_nan=True), key
        assert np.array_equal(z["q_after"], g["both_q_after"][rank::world], equal_nan=True)
        assert np.array_equal(z["rep_q"], g["both_q_after"][rank::world], equal_nan=True)
        assert np.array_equal(z["inds"], g["both_inds"][rank::world])
        assert np.array_equal(z["rep_inds"], g["both_inds"][rank::world])
        assert int(z["nc"]) == int(g["nc"])
        # float32 coordinates are upcast, not reinterpreted: same NaN pattern, values close to the float64 answer
        ok = ~np.isnan(z["f32"]).any(axis=1) & ~np.isnan(ref[rank::world]).any(axis=1)
        assert ok.sum() > 10 and np.abs(z["f32"][ok][:, :4] - ref[rank::world][ok][:, :4]).max() < 1e-3
