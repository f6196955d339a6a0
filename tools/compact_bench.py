"""Throughput of the wavefront compaction kernel (k_compact_u32 + k_zero_tail_u32) against the HBM copy peak.
    python tools/compact_bench.py [iters]
Bytes per element that the contract forces: 4 B value read + 1 B keep flag read + 4 B written per element either way (kept values
in front, zeros behind: the reference's output is zero past the new size) = 9 B; the look-back status words add 8 B per 2048
elements.  Prints one JSON line per (n, keep fraction): device ms (CUDA events, L2 flushed before every launch), GB/s, fraction of
the measured HBM peak; checks the result against numpy once per case."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rendiation_b200 import api  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
s = api.NaiveSahBVHSystem()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n in (2_073_600, 132_710_400):
    vals = torch.randint(0, 2 ** 31 - 1, (n,), dtype=torch.int32, device="cuda")
    out = torch.empty_like(vals)
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    for frac in (0.3, 1.0):
        keep = (torch.rand(n, device="cuda") < frac).to(torch.uint8) if frac < 1.0 else torch.ones(n, dtype=torch.uint8, device="cuda")
        s.compact_u32_device(vals.data_ptr(), keep.data_ptr(), n, out.data_ptr(), cnt.data_ptr(), stream=st)
        torch.cuda.synchronize()
        k = int(cnt.item())
        want = vals[keep.bool()]
        ok = k == want.numel() and bool(torch.equal(out[:k], want)) and bool((out[k:] == 0).all())
        ms = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s.compact_u32_device(vals.data_ptr(), keep.data_ptr(), n, out.data_ptr(), cnt.data_ptr(), stream=st)
            e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms = float(np.median(ms))
        nbytes = 9.0 * n + 8.0 * (n / 2048)
        print(json.dumps({"kernel": ("k_compact_count + k_compact_scan_groups + k_compact_scatter" if n >= (1 << 22) else "k_compact_u32 + k_zero_tail_u32"), "n": n, "keep_fraction": frac, "kept": k, "ms_median": ms, "iters": iters,
                          "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "hbm_peak_gbs": peak, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak,
                          "elements_per_s": n / ms * 1e3, "matches_numpy": ok}))
