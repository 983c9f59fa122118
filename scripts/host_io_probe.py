"""Wall clock of the C++ host mirror's O(n) host paths on this machine: DeepArcManager::read / write and
the flatten() gather of solve(), parallel (default) vs DEEPARC_SERIAL_IO=1, on a 1M-observation file."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from deeparc_sfm_b200 import synthetic
p = synthetic.bal_like(n_cam=340, n_pts=200_000, obs_per_point=5, window=50, name="bal1m")
path = "/tmp/bal1m.deeparc"
synthetic.write_deeparc(p, path)
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deeparc-sfm_b200/lib/libdeeparc_host.so"))
lib.dam_manager_read.restype = ctypes.c_void_p
lib.dam_manager_read.argtypes = [ctypes.c_char_p]
lib.dam_manager_write.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
lib.dam_manager_free.argtypes = [ctypes.c_void_p]
lib.dam_manager_export.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 15
n = p.n_obs
print(f"{n} observations, {os.path.getsize(path) / 1e6:.0f} MB, {os.cpu_count()} logical CPUs")
for mode in ("1", "0", "0"):
    os.environ["DEEPARC_SERIAL_IO"] = mode
    t = time.perf_counter(); h = lib.dam_manager_read(path.encode()); tr = time.perf_counter() - t
    xy = np.zeros(2 * n); a = [np.zeros(n, np.int32) for _ in range(4)]
    t = time.perf_counter()
    lib.dam_manager_export(h, xy.ctypes.data, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data, *([None] * 10))
    tf = time.perf_counter() - t
    t = time.perf_counter(); lib.dam_manager_write(h, b"/tmp/bal1m_out.deeparc"); tw = time.perf_counter() - t
    t = time.perf_counter(); lib.dam_manager_free(h); td = time.perf_counter() - t
    print(("serial  " if mode == "1" else "parallel"), f"read {1e3*tr:7.1f} ms  flatten {1e3*tf:6.1f} ms  write {1e3*tw:7.1f} ms  free {1e3*td:6.1f} ms")
