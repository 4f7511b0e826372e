"""distCUDA2 (simple-knn) timing: this library against the reference's own CUDA build (oracle/_ref/ref_knn_C.so) on the same
points, plus bit-exactness of the result:  python scripts/time_knn.py [P ...]   (default 1M and 6M points, uniform in a cube)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402

Pk = H.pkg()
ref = H.ref_knn()
sizes = [int(a) for a in sys.argv[1:]] or [1_000_000, 6_000_000]


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for P in sizes:
    pts = (torch.rand(P, 3, generator=torch.Generator().manual_seed(P % 1000)) * 12.0 - 6.0).cuda()
    ours = Pk.distCUDA2(pts)
    out = {"P": P, "ours_ms": round(timeit(lambda: Pk.distCUDA2(pts)), 3)}
    if ref is not None:
        want = ref.distCUDA2(pts)
        out["reference_ms"] = round(timeit(lambda: ref.distCUDA2(pts)), 3)
        out["speedup"] = round(out["reference_ms"] / out["ours_ms"], 2)
        out["bit_exact"] = bool(torch.equal(ours, want))
    print(json.dumps(out))
