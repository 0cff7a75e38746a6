import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_stitcher_b200 import _lib
out_d = torch.rand(9012, 9012, device="cuda")
out_h = np.zeros((9012, 9012), np.float32)
st2 = torch.cuda.Stream(); sp2 = ctypes.c_void_p(st2.cuda_stream)
torch.cuda.synchronize()
for i in range(4):
    t0 = time.perf_counter(); _lib.copy_d2h(out_h, out_d, sp2); print("whole", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
for i in range(2):
    t0 = time.perf_counter(); _lib.copy_d2h(out_h[:2048], out_d[:2048], sp2); print("band", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
for i in range(2):
    t0 = time.perf_counter(); _lib.copy_d2h(out_h[:2048, :8192], out_d[:2048, :8192], sp2); print("band-pitched", (time.perf_counter() - t0) * 1e3, file=sys.stderr)
assert np.array_equal(out_h, out_d.cpu().numpy())
