"""Worker of the 2-GPU test (launched by torch.distributed.run, one rank per GPU): the
point-sharded solve must reproduce the single-GPU solve to reduction-order round-off — with the
camera-space vectors combined by ncclAllReduce (few camera blocks: the per-camera sum is split)
and by the fused PCG tail that exchanges them through the NVLink peer windows ("bal320")."""
import ctypes
import os
import sys

sys.path.insert(0, os.environ.get("REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist

from deeparc_sfm_b200 import capi, synthetic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = capi.load_library()
buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    raw = ctypes.create_string_buffer(128)
    assert lib.dba_nccl_unique_id(raw) == 0
    buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).cuda()
dist.broadcast(buf, 0)
uid = bytes(buf.cpu().numpy().tobytes())

eng = capi.Engine(device=local, rank=rank, world_size=world, nccl_unique_id=uid)  # one communicator per unique id
fused_before = 0
for name, p in (("bal", synthetic.bal_like(n_cam=60, n_pts=6000, window=12, seed=91)),
                ("rig", synthetic.arc_rig(n_arc=4, n_ring=5, n_pts=3000, obs_per_point=8, seed=92)),
                ("rig_dense", synthetic.arc_rig(n_arc=4, n_ring=5, n_pts=3000, obs_per_point=8, seed=92)),
                ("bal320", synthetic.bal_like(n_cam=320, n_pts=24000, window=20, seed=93))):
    # "rig_dense": the explicit reduced system, summed over the ranks by one ncclAllReduce, factorised on every rank
    opts = capi.make_options(max_num_iterations=5, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0,
                             linear_solver=capi.DBA_LS_DENSE if name == "rig_dense" else capi.DBA_LS_PCG,
                             pcg_rel_tolerance=1e-13, pcg_max_iterations=3000)
    eng.problem_set(p)
    c0 = eng.eval(residuals=False)["cost"]
    s = eng.solve(opts)
    x = eng.params_get()
    launched = {k["name"]: k["launches"] for k in eng.kernel_stats()}
    if name == "bal320" and os.environ.get("DBA_P2P", "1") != "0":
        # >= 296 camera blocks: one fused launch per PCG iteration, no ncclAllReduce on the PCG path
        assert launched.get("spmv_mf_pcg", 0) + launched.get("pcg_fused", 0) > 0 and launched.get("partials_to_q", 0) == fused_before, launched
    fused_before = launched.get("partials_to_q", 0)
    if rank == 0:
        one = capi.Engine(device=local)
        one.problem_set(p)
        c1 = one.eval(residuals=False)["cost"]
        s1 = one.solve(opts)
        x1 = one.params_get()
        one.close()
        assert abs(c0 - c1) <= 1e-12 * c1, (c0, c1)
        np.testing.assert_allclose(s.trace("cost"), s1.trace("cost"), rtol=1e-9)
        assert np.array_equal(s.trace("step_is_successful"), s1.trace("step_is_successful"))
        for k in x:
            assert np.max(np.abs(x[k] - x1[k])) <= 1e-8 * max(np.max(np.abs(x1[k])), 1e-300), (name, k)
        print(name, "sharded == single:", s.final_cost, s1.final_cost, flush=True)
    dist.barrier()
eng.close()
if rank == 0:
    print("MGPU OK", flush=True)
dist.destroy_process_group()
