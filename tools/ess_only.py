import sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import torch, genpf_b200 as g
L, lib = g._lib, g.load()
n = 1 << 24
lw = torch.randn(n, dtype=torch.float64, device="cuda")
ess = C.c_double()
for _ in range(6):
    L.check(lib.genpf_ess(lw.data_ptr(), n, L.DEVICE_PTRS, C.byref(ess)))
print(ess.value)
