#!/usr/bin/env python
"""Radix sort alone: N random keys of `bits` bits, CUDA-event time per sort (the variant comes from the environment:
HC_SORT_LOOKBACK = 1|2|4|8, HC_SORT_MINB = 3|4)."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from hichap_master_b200 import _abi, kernels  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=500_000_000)
ap.add_argument("--bits", type=int, default=38)
ap.add_argument("--begin", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
keys = torch.randint(0, 1 << a.bits, (a.n,), dtype=torch.int64, device=dev, generator=g) << a.begin
src, tmp = torch.empty_like(keys), torch.empty_like(keys)
work = torch.empty(int(_abi.lib().hc_sort_work_bytes(a.n)), dtype=torch.uint8, device=dev)
flag = C.c_int32(0)
ms = []
for rep in range(a.reps + 1):
    src.copy_(keys)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _abi.check(_abi.lib().hc_sort_keys_u64(kernels.ptr(src), kernels.ptr(tmp), a.n, a.begin, a.begin + a.bits, kernels.ptr(work),
                                           C.byref(flag), kernels.stream_ptr()), "sort")
    e1.record()
    torch.cuda.synchronize()
    if rep:
        ms.append(e0.elapsed_time(e1))
out = tmp if flag.value else src
ok = bool((out[1:] >= out[:-1]).all().item())
passes = (a.bits + 7) // 8
m = min(ms)
print(json.dumps({"n": a.n, "bits": a.bits, "passes": passes, "ms": ms, "sorted": ok,
                  "GBps": (8.0 + 16.0 * passes) * a.n / (m * 1e6), "Gkeys_per_s": a.n / (m * 1e6),
                  "lookback": os.environ.get("HC_SORT_LOOKBACK", "8"), "minb": os.environ.get("HC_SORT_MINB", "4")}))
