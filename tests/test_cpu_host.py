"""CPU tests (no GPU) of the boundary and the host side: the C ABI library loads and exports
every symbol include/deeparc_ba.h declares, fails loudly without a device, the C++ mirror of the
reference's DeepArcManager reads / writes exactly what the reference's own code does, and the
multi-GPU sharding plan is consistent across ranks (world_size-2 gloo)."""
import filecmp
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from deeparc_sfm_b200 import capi, synthetic
from tests import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    return capi.load_library().dba_device_count() <= 0


# ----------------------------------------------------------------------------------- ABI
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "deeparc_ba.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(dba_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 17, declared
    lib = capi.load_library()
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.dba_abi_version() == 2


def test_struct_layouts_match_header_sizes():
    # sizes computed by the C compiler from the header, compared with the ctypes mirrors
    src = '#include <stdio.h>\n#include "deeparc_ba.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(dba_config), sizeof(dba_problem), sizeof(dba_solve_options), sizeof(dba_iteration), sizeof(dba_summary), sizeof(dba_kernel_stat));return 0;}'
    d = os.path.join(ROOT, "tests", "_build")
    os.makedirs(d, exist_ok=True)
    c = os.path.join(d, "abi_sizes.c")
    open(c, "w").write(src)
    exe = os.path.join(d, "abi_sizes")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
    sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    import ctypes as C
    mine = [C.sizeof(t) for t in (capi.DbaConfig, capi.DbaProblem, capi.DbaSolveOptions, capi.DbaIteration,
                                  capi.DbaSummary, capi.DbaKernelStat)]
    assert sizes == mine, (sizes, mine)


def test_default_options_match_python_mirror():
    lib = capi.load_library()
    o = capi.DbaSolveOptions()
    lib.dba_solve_options_default(o)
    ref = capi.default_options_struct()
    for name, _ in capi.DbaSolveOptions._fields_:
        assert getattr(o, name) == getattr(ref, name), name


def test_no_cpu_fallback_without_device():
    if not _no_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.EngineError) as e:
        capi.Engine(device=0)
    assert e.value.status == capi.DBA_ERR_NO_DEVICE
    assert "no CPU path" in str(e.value)


def test_sfm_driver_fails_loudly_without_device(tmp_path):
    if not _no_gpu():
        pytest.skip("a GPU is present")
    p = synthetic.arc_rig(n_arc=2, n_ring=3, n_pts=20, obs_per_point=4)
    f = str(tmp_path / "in.deeparc")
    synthetic.write_deeparc(p, f)
    exe = os.path.join(ROOT, "deeparc-sfm_b200", "bin", "sfm")
    r = subprocess.run([exe, "--input", f, "--output", str(tmp_path / "out.deeparc"), "--ply-init", str(tmp_path / "i.ply"),
                        "--ply-adjust", str(tmp_path / "a_"), "--ply-clear", str(tmp_path / "c.ply")],
                       capture_output=True, text=True)
    assert r.returncode == 1
    assert "cannot create the GPU engine" in r.stderr


def test_sfm_driver_missing_input_reports_like_reference(tmp_path):
    exe = os.path.join(ROOT, "deeparc-sfm_b200", "bin", "sfm")
    r = subprocess.run([exe, "--input", str(tmp_path / "nope.deeparc")], capture_output=True, text=True)
    assert r.returncode == 1
    assert "Cannot read" in r.stdout and "Cannot read input file" in r.stderr  # DeepArcManager.cc:28-31


# ---------------------------------------------------------------------- host mirror vs ref
CASES = [("rig_aa", dict(kind="rig", fmt=3)), ("rig_matrix", dict(kind="rig", fmt=9)), ("rig_quat", dict(kind="rig", fmt=4)),
         ("plain", dict(kind="bal", fmt=3))]


def _make(kind):
    if kind == "rig":
        return synthetic.arc_rig(n_arc=3, n_ring=4, n_pts=120, obs_per_point=6, seed=51)
    return synthetic.bal_like(n_cam=12, n_pts=100, window=6, seed=52, free_intrinsics=0)


@pytest.mark.parametrize("name,cfg", CASES)
def test_host_mirror_reads_and_writes_like_the_reference(reference, tmp_path, name, cfg):
    H = oracle_lib.HostMirror()
    p = _make(cfg["kind"])
    f = str(tmp_path / (name + ".deeparc"))
    synthetic.write_deeparc(p, f, rotation_format=cfg["fmt"], center_override=(923.5, 1223.5))
    hr, hh = reference.read(f), H.read(f)
    assert reference.counts(hr) == H.counts(hh)
    assert reference.is_shared(hr) == H.is_shared(hh)
    a, b = reference.export(hr), H.export(hh)
    for k in ("obs_xy", "obs_pt", "obs_pose_a", "obs_pose_b", "obs_intr", "pts", "ext_rot", "ext_trans", "intr_center",
              "intr_focal", "intr_dist", "intr_nf", "intr_nd", "ext_const", "pts_rgb"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert np.array_equal(b.intr_center[0], [923.0, 1223.0])  # Intrinsic::center(int,int) truncation
    reference.write(hr, f + ".ref"), H.write(hh, f + ".host")
    assert filecmp.cmp(f + ".ref", f + ".host", shallow=False)
    reference.write_ply(hr, f + ".ref.ply"), H.write_ply(hh, f + ".host.ply")
    assert filecmp.cmp(f + ".ref.ply", f + ".host.ply", shallow=False)
    assert np.array_equal(reference.camera_centers(hr), H.camera_centers(hh))
    reference.free(hr), H.free(hh)


@pytest.mark.parametrize("kind", ["rig", "bal"])
def test_binary_side_format_is_lossless(tmp_path, kind):
    """the text format keeps six decimals (reference src/DeepArcManager.cc:428), so a
    text round trip moves every parameter; the binary side format (writeBinary, recognised by read()
    through its magic, loaded from the mmap'ed image) reproduces the scene bit for bit, and writing
    text from either copy gives the same bytes."""
    H = oracle_lib.HostMirror()
    p = _make(kind)
    f = str(tmp_path / "a.deeparc")
    synthetic.write_deeparc(p, f)
    h = H.read(f)
    a = H.export(h)
    H.write_binary(h, f + "b")
    H.write(h, f + ".txt1")
    hb = H.read(f + "b")
    b = H.export(hb)
    assert H.counts(h) == H.counts(hb) and H.is_shared(h) == H.is_shared(hb)
    for k in ("obs_xy", "obs_pt", "obs_pose_a", "obs_pose_b", "obs_intr", "pts", "ext_rot", "ext_trans", "intr_center",
              "intr_focal", "intr_dist", "intr_nf", "intr_nd", "ext_const", "pts_rgb"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    H.write(hb, f + ".txt2")
    assert filecmp.cmp(f + ".txt1", f + ".txt2", shallow=False)
    H.write_binary(hb, f + "b2")
    assert filecmp.cmp(f + "b", f + "b2", shallow=False)
    # the text round trip is lossy, which is what the side format is for
    ht = H.read(f + ".txt1")
    t = H.export(ht)
    assert not np.array_equal(t.pts, a.pts) and np.allclose(t.pts, a.pts, atol=1e-6)
    assert np.array_equal(H.camera_centers(h), H.camera_centers(hb))
    for x in (h, hb, ht):
        H.free(x)
    # a truncated binary file is refused, not half-loaded
    raw = open(f + "b", "rb").read()
    open(f + "b3", "wb").write(raw[: len(raw) // 2])
    with pytest.raises(Exception):
        H.read(f + "b3")


def test_host_mirror_standalone_roundtrip(tmp_path):
    """Without the reference build: read -> write -> read reproduces the scene to the 6 decimals
    of the text format; ids, modes and the gauge mask survive."""
    H = oracle_lib.HostMirror()
    p = synthetic.arc_rig(n_arc=3, n_ring=3, n_pts=80, obs_per_point=5, seed=53)
    f = str(tmp_path / "a.deeparc")
    synthetic.write_deeparc(p, f)
    h = H.read(f)
    a = H.export(h)
    for k in ("obs_pt", "obs_pose_a", "obs_pose_b", "obs_intr", "ext_const"):
        assert np.array_equal(getattr(a, k), getattr(p, k)), k
    assert np.array_equal(a.pts, p.pts) and np.array_equal(a.ext_rot, p.ext_rot)
    H.write(h, f + ".out")
    h2 = H.read(f + ".out")
    b = H.export(h2)
    assert np.array_equal(a.obs_pose_a, b.obs_pose_a) and np.array_equal(a.obs_pose_b, b.obs_pose_b)
    np.testing.assert_allclose(b.pts, a.pts, atol=5.1e-7)
    np.testing.assert_allclose(b.ext_trans, a.ext_trans, atol=5.1e-7)
    text = synthetic.read_deeparc_text(f + ".out")
    assert text["version"] == 0.01 and all(len(r) == 3 for _, r in text["ext"])
    H.free(h), H.free(h2)


def test_host_mirror_bad_point_id_raises_out_of_range(tmp_path):
    H = oracle_lib.HostMirror()
    p = synthetic.arc_rig(n_arc=2, n_ring=2, n_pts=10, obs_per_point=3, seed=54)
    p.obs_pt = p.obs_pt.copy()
    p.obs_pt[0] = 999
    f = str(tmp_path / "bad.deeparc")
    synthetic.write_deeparc(p, f)
    with pytest.raises(IOError):
        H.read(f)
    assert "vector" in H.last_error() or "range" in H.last_error()


# --------------------------------------------------------------- parallel loader / writer
def test_fixed6_formatter_is_byte_identical_to_printf():
    """The writer formats "%.6f" with an integer fast path and falls back to snprintf near rounding
    ties and for large values: every output must be what printf gives (DeepArcManager.cc:428)."""
    import ctypes
    H = oracle_lib.HostMirror()
    f = H.lib.dam_format_fixed6
    f.argtypes = [ctypes.c_double, ctypes.c_char_p]
    f.restype = ctypes.c_int
    buf = ctypes.create_string_buffer(512)
    rng = np.random.default_rng(7)
    vals = [0.0, -0.0, 1e-7, -1e-7, 4.9e-7, 5e-7, -5e-7, 5.1e-7, 0.0078125, 0.0000005, 1.0000005, 2.5e-6, 3.5e-6,
            0.1234565, 0.1234575, 999999.9999994, 999999.9999996, 1e6, -1e6, 123456789.123456789, 1e15, 1e22, -3e300,
            4949.234294, 923.0, 0.5, 1 / 3, 2 / 3, 1e-300, 5e-324, float("inf"), float("-inf")]
    vals += [(k + 0.5) / 1e6 for k in range(0, 4000, 7)] + [-(k + 0.5) / 1e6 for k in range(1, 400, 3)]
    vals += [k / 64.0 + j * 2.0 ** -20 for k in range(40) for j in range(4)]  # dyadic: exact ties possible
    vals += list(rng.standard_normal(4000) * 10.0 ** rng.integers(-8, 9, 4000))
    vals += list(np.nextafter(np.array([(k + 0.5) / 1e6 for k in range(50)]), 1e9))
    vals += list(np.nextafter(np.array([(k + 0.5) / 1e6 for k in range(50)]), -1e9))
    for v in vals:
        v = float(v)
        n = f(v, buf)
        assert buf.value[:n].decode() == "%.6f" % v, (v, buf.value, "%.6f" % v)


def test_parallel_loader_and_writer_match_the_serial_path(tmp_path, monkeypatch):
    """DeepArcManager::read parses, allocates and links the observation and point sections with all
    host cores; DEEPARC_SERIAL_IO=1 selects the single-threaded tokenizer.  Same scene, same bytes."""
    H = oracle_lib.HostMirror()
    outs = {}
    for kind, p in (("rig", synthetic.arc_rig(n_arc=4, n_ring=6, n_pts=3000, obs_per_point=7, seed=61)),
                    ("bal", synthetic.bal_like(n_cam=40, n_pts=20000, window=10, seed=62, free_intrinsics=0))):
        f = str(tmp_path / (kind + ".deeparc"))
        synthetic.write_deeparc(p, f)
        for mode in ("1", "0"):
            monkeypatch.setenv("DEEPARC_SERIAL_IO", mode)
            h = H.read(f)
            outs[mode] = (H.counts(h), H.export(h))
            H.write(h, f + ".out" + mode)
            H.write_ply(h, f + ".ply" + mode)
            H.free(h)
        assert outs["0"][0] == outs["1"][0]
        for k in ("obs_xy", "obs_pt", "obs_pose_a", "obs_pose_b", "obs_intr", "pts", "pts_rgb", "ext_rot", "ext_trans",
                  "intr_center", "intr_focal", "intr_dist", "intr_nf", "intr_nd", "ext_const"):
            assert np.array_equal(getattr(outs["0"][1], k), getattr(outs["1"][1], k)), (kind, k)
        assert np.array_equal(outs["0"][1].obs_pt, p.obs_pt)
        assert filecmp.cmp(f + ".out0", f + ".out1", shallow=False)
        assert filecmp.cmp(f + ".ply0", f + ".ply1", shallow=False)


def test_parallel_loader_handles_free_form_whitespace(tmp_path, monkeypatch):
    """The format is a token stream (the reference reads it with operator>>): CRLF line ends, tabs, runs of
    blanks, several records per line and records split across lines load the same as the canonical layout."""
    H = oracle_lib.HostMirror()
    p = synthetic.arc_rig(n_arc=3, n_ring=4, n_pts=400, obs_per_point=6, seed=64)
    f = str(tmp_path / "canon.deeparc")
    synthetic.write_deeparc(p, f)
    tokens = open(f).read().split()
    rng = np.random.default_rng(11)
    seps = [" ", "  ", "\t", "\r\n", "\n", " \n\t "]
    messy = "".join(tok + seps[int(rng.integers(len(seps)))] for tok in tokens)
    g = str(tmp_path / "messy.deeparc")
    open(g, "w", newline="").write("\r\n \t" + messy)
    monkeypatch.setenv("DEEPARC_SERIAL_IO", "0")
    ha, hb = H.read(f), H.read(g)
    a, b = H.export(ha), H.export(hb)
    for k in ("obs_xy", "obs_pt", "obs_pose_a", "obs_pose_b", "obs_intr", "pts", "pts_rgb", "ext_rot", "ext_trans",
              "intr_center", "intr_focal", "intr_dist", "ext_const"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    H.write(ha, f + ".out"), H.write(hb, g + ".out")
    assert filecmp.cmp(f + ".out", g + ".out", shallow=False)
    H.free(ha), H.free(hb)


def test_parallel_loader_falls_back_on_tokens_it_does_not_take(tmp_path, monkeypatch):
    """The parallel parser is strict (a token must be one number); a file with e.g. a hexadecimal
    float or a truncated tail is re-read by the serial tokenizer, which has the reference's
    istream-like prefix semantics.  Both routes give the same scene."""
    H = oracle_lib.HostMirror()
    p = synthetic.bal_like(n_cam=6, n_pts=50, window=4, seed=63, free_intrinsics=0)
    f = str(tmp_path / "odd.deeparc")
    synthetic.write_deeparc(p, f)
    text = open(f).read().split("\n")
    rec = text[2].split()          # first observation line: a r pid x y
    rec[3] = float(rec[3]).hex()   # strtod reads it, std::from_chars(general) does not
    text[2] = " ".join(rec)
    open(f, "w").write("\n".join(text))
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("DEEPARC_SERIAL_IO", mode)
        h = H.read(f)
        res[mode] = H.export(h)
        H.free(h)
    assert np.array_equal(res["0"].obs_xy, res["1"].obs_xy) and np.array_equal(res["0"].obs_pt, res["1"].obs_pt)
    assert abs(res["0"].obs_xy[0, 0] - p.obs_xy[0, 0]) < 1e-6
    # truncated file: fewer tokens than the header promises -> the serial path's zeros, no crash
    open(f, "w").write("\n".join(text[: len(text) // 2]))
    for mode in ("1", "0"):
        monkeypatch.setenv("DEEPARC_SERIAL_IO", mode)
        try:
            h = H.read(f)
            H.free(h)
        except IOError:
            pass  # ids of the missing sections may be out of range, as in the reference


# ------------------------------------------------------------------------- sharding plan
def test_shard_plan_covers_and_balances():
    p = synthetic.bal_like(n_cam=40, n_pts=5000, window=10, seed=61)
    for world in (1, 2, 3, 8):
        pt_begin, obs_count = capi.shard_plan(p, world)
        assert pt_begin[0] == 0 and pt_begin[-1] == p.n_pts and np.all(np.diff(pt_begin) >= 0)
        assert obs_count.sum() == p.n_obs
        assert obs_count.max() - obs_count.min() <= 2 * 5  # within a couple of tracks
    # ragged: empty problem and a single huge track
    e = synthetic.bal_like(n_cam=3, n_pts=4, obs_per_point=3, window=3, seed=62)
    e.obs_pt = np.zeros_like(e.obs_pt)
    pt_begin, obs_count = capi.shard_plan(e, 4)
    assert obs_count.sum() == e.n_obs and pt_begin[-1] == e.n_pts


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["REPO_ROOT"])
import numpy as np, torch, torch.distributed as dist
from deeparc_sfm_b200 import capi, synthetic
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
p = synthetic.bal_like(n_cam=30, n_pts=2000, window=10, seed=63)      # every rank builds the full problem
pt_begin, obs_count = capi.shard_plan(p, world)
lo, hi = int(pt_begin[rank]), int(pt_begin[rank + 1])
mine = (p.obs_pt >= lo) & (p.obs_pt < hi)
assert int(mine.sum()) == int(obs_count[rank])
# every observation is owned by exactly one rank
owned = torch.from_numpy(mine.astype(np.int64)); dist.all_reduce(owned); assert bool((owned == 1).all())
# a camera-space vector (observations per camera) summed over shards == the global one: the
# collective the engine runs (allreduce of per-camera accumulators), with gloo standing in for NCCL
local = torch.from_numpy(np.bincount(p.obs_pose_a[mine], minlength=p.n_ext).astype(np.float64))
dist.all_reduce(local)
assert np.array_equal(local.numpy(), np.bincount(p.obs_pose_a, minlength=p.n_ext).astype(np.float64))
# scalar reductions (cost): sum of shard partials == global
from tests import oracle_lib
res = oracle_lib.Oracle().eval(p)["residuals"]
part = torch.tensor([0.5 * float((res[mine] ** 2).sum())], dtype=torch.float64); dist.all_reduce(part)
assert abs(part.item() - 0.5 * float((res ** 2).sum())) <= 1e-12 * part.item()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_shard_plan_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


# ---------------------------------------------------------------------------- generators
def test_generators_follow_the_reference_aliasing_rules():
    p = synthetic.arc_rig(n_arc=4, n_ring=5, n_pts=300, obs_per_point=9, seed=71)
    arc, ring = p.obs_col0, p.obs_col1
    A = p.n_arc
    ring_slot = np.where(ring == 0, 0, ring + A - 1)
    assert np.array_equal(p.obs_intr, arc)
    assert np.array_equal(p.obs_pose_a, np.where((arc == 0) & (ring != 0), ring_slot, arc))
    assert np.array_equal(p.obs_pose_b, np.where((arc != 0) & (ring != 0), ring_slot, -1))
    assert p.ext_const[0] == 1 and p.ext_const[1:].sum() == 0
    assert p.n_ext == A + p.n_ring - 1
    r = synthetic.project(p, pts=p.truth["pts"], ext_rot=p.truth["ext_rot"], ext_trans=p.truth["ext_trans"]) - p.obs_xy
    assert 0.3 < r.std() < 0.7  # 0.5 px observation noise around the ground truth


# ------------------------------------------------------------------------- bench contract
def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU oracle timed on the host cores) prints exactly one JSON line
    carrying the contract keys; it needs no GPU."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                        "--steps", "1", "--cpu-budget", "120"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "lm_iters_per_sec" and d["unit"] == "LM iterations/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # 200 cameras x 9 = 1800 reduced unknowns: the arm runs the algorithm the GPU arm times (implicit Schur
    # PCG, same K) and says so; the DENSE_SCHUR time of one iteration is reported beside it
    assert "PCG" in d["config"]["linear_solver"] and "PCG" in d["cpu_baseline"]["sample"]
    assert d["cpu_baseline_dense_schur"]["value"] > 0 and "DENSE_SCHUR" in d["cpu_baseline_dense_schur"]["sample"]


def test_bench_fails_loudly_without_a_device():
    if not _no_gpu():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "small", "--steps", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
