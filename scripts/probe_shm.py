import os, time, numpy as np, torch
from multiprocessing import shared_memory, resource_tracker
st = os.statvfs("/dev/shm"); print("shm free GB", st.f_bavail * st.f_frsize / 1e9)
size = 400 << 20
shm = shared_memory.SharedMemory(create=True, size=size)
a = np.ndarray((size // 8,), dtype=np.float64, buffer=shm.buf)
t0 = time.perf_counter(); a[:] = 0; print("touch s", time.perf_counter() - t0)
rt = torch.cuda.cudart()
t0 = time.perf_counter(); rc = rt.cudaHostRegister(a.ctypes.data, size, 0); print("register rc", rc, "s", time.perf_counter() - t0)
t = torch.from_numpy(a)
print("is_pinned", t.is_pinned())
d = torch.randn(size // 8, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); t.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("d2h GB/s", size / dt / 1e9)
print("equal", bool((t == d.cpu()).all()))
rc = rt.cudaHostUnregister(a.ctypes.data); print("unregister", rc)
del t, a
shm.close(); shm.unlink()
