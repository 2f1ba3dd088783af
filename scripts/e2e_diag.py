import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from distributed_sddmm_b200 import driver as D, lib
L = lib()
D.world_init("self")
h = torch.full((1 << 20, 128), 0.001, dtype=torch.float64).pin_memory()
print("pinned:", h.is_pinned())
import ctypes as C
d = C.c_void_p()
D.check(L.hnhd_dense_create(1 << 20, 128, 0.0, C.byref(d)))
for i in range(3):
    t0 = time.perf_counter(); D.check(L.hnhd_dense_from_host(d, h.data_ptr())); t1 = time.perf_counter()
    D.check(L.hnhd_dense_to_host(d, h.data_ptr())); t2 = time.perf_counter()
    print(f"h2d {1.0737/(t1-t0):.1f} GB/s  d2h {1.0737/(t2-t1):.1f} GB/s")
g = torch.empty((1 << 20, 128), dtype=torch.float64, device="cuda")
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); g.copy_(h, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"torch h2d {1.0737/(t1-t0):.1f} GB/s")
