"""hook C end to end on C2 (the bench's e2e step), repeated, with the transfer counters."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from multiview_stitcher_b200 import geometry, synthetic
from multiview_stitcher_b200.batch import BatchFuser, block_geometry
views, stage, true = synthetic.make_grid(bench.GRID, bench.TILE, bench.OVERLAP, np.float32, jitter=2, seed=1, subpixel=True)
bbs = [v.bb() for v in views]
osp = geometry.union_stack_props(bbs, true, bbs[0]["spacing"])
host_tiles = [v.tensor.cpu().numpy() for v in views]
msims = [{"data": d, "origin": v.origin, "spacing": v.spacing, "transforms": {"reg": p}} for d, v, p in zip(host_tiles, views, true)]
out_host = np.zeros(tuple(int(osp["shape"][d]) for d in "yx"), dtype=np.float32)
chunksize = {"y": int(os.environ.get("CHUNK", 2048)), "x": int(os.environ.get("CHUNK", 2048))}
fuse_chunk = bench._fake_partial(msims, osp, chunksize, out_host)
ids = sorted(block_geometry(osp, chunksize))
bf = BatchFuser()
ts = []
for i in range(8):
    bf.reset(); t0 = time.perf_counter(); bf(fuse_chunk, ids); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print(os.environ.get("MVS_COPY_THREADS"), chunksize["y"], " ".join(f"{t:.1f}" for t in ts))
