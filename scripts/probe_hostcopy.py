"""Where hook C's end-to-end time goes: staging copies vs DMA vs the host's own memcpy rate."""
import ctypes, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiview_stitcher_b200 import _lib

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

res = {"cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
tiles = [np.random.rand(2048, 2048).astype(np.float32) for _ in range(25)]
dev = [torch.empty(2048, 2048, device="cuda") for _ in range(25)]
out_d = torch.rand(9012, 9012, device="cuda")
out_h = np.zeros((9012, 9012), np.float32)
st = torch.cuda.Stream(); st2 = torch.cuda.Stream()
sp = ctypes.c_void_p(st.cuda_stream); sp2 = ctypes.c_void_p(st2.cuda_stream)
res["h2d_staged_ms"] = t(lambda: [_lib.copy_h2d(d, h, sp) for d, h in zip(dev, tiles)])
res["d2h_staged_ms"] = t(lambda: _lib.copy_d2h(out_h, out_d, sp2))
def blocks():
    for y in range(0, 9012, 2048):
        for x in range(0, 9012, 2048):
            _lib.copy_d2h(out_h[y:y+2048, x:x+2048], out_d[y:y+2048, x:x+2048], sp2)
res["d2h_staged_blocks_ms"] = t(blocks)
import threading
def both():
    th = threading.Thread(target=blocks); th.start()
    [_lib.copy_h2d(d, h, sp) for d, h in zip(dev, tiles)]
    th.join()
res["both_staged_ms"] = t(both)
pin = [torch.from_numpy(a).pin_memory() for a in tiles]
pout = torch.empty(9012, 9012).pin_memory()
def pin_h2d():
    with torch.cuda.stream(st):
        for d, h in zip(dev, pin): d.copy_(h, non_blocking=True)
def pin_d2h():
    with torch.cuda.stream(st2): pout.copy_(out_d, non_blocking=True)
res["h2d_pinned_ms"] = t(pin_h2d)
res["d2h_pinned_ms"] = t(pin_d2h)
res["both_pinned_ms"] = t(lambda: (pin_h2d(), pin_d2h()))
big = np.concatenate([a.ravel() for a in tiles]); dst = np.empty_like(big)
res["numpy_memcpy_419MB_ms"] = t(lambda: np.copyto(dst, big), 3)
res["pageable_torch_h2d_ms"] = t(lambda: [d.copy_(torch.from_numpy(h)) for d, h in zip(dev, tiles)], 3)
print(json.dumps(res, indent=1))
json.dump(res, open("gpurun_out/probe_hostcopy.json", "w"), indent=1)
