"""torchrun worker: wall-clock split of the end-to-end path (problem_set / solve / params_get) per rank."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from deeparc_sfm_b200 import capi
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
eng = capi.Engine(device=local, rank=rank, world_size=world, nccl_unique_id=bytes(buf.cpu().numpy().tobytes()))
p = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else "bal5m")
for i in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter(); eng.problem_set(p); t1 = time.perf_counter()
    s = eng.solve(bench.solve_options(capi, 10, 20)); t2 = time.perf_counter()
    out = eng.params_get(); t3 = time.perf_counter()
    print(f"[rank {rank}] problem_set {1e3*(t1-t0):.1f} ms  solve {1e3*(t2-t1):.1f} ms (device loop {1e3*s.loop_device_time_in_seconds:.1f})  get {1e3*(t3-t2):.1f} ms", file=sys.stderr, flush=True)
eng.close()
dist.destroy_process_group()
