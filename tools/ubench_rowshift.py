"""Probe: tcgen05.mma SWIZZLE_128B K-major descriptors whose start is offset by whole rows (see ubench_sm100.cu)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ensemble_svs_with_interactions_b200.csrc.build import build_ubench  # noqa: E402

l = C.CDLL(build_ubench())   # tools/libsvsk_ubench.so — the micro-benchmarks are not part of the product libsvsk.so
l.svsk_ubench_rowshift.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
l.svsk_last_error.restype = C.c_char_p
rows = 160
win = torch.randn(rows, 64, device="cuda").to(torch.bfloat16)
for mode in (0, 1):
    for r0 in list(range(0, 18)) + [23, 32]:
        out = torch.full((128, 64), float("nan"), device="cuda")
        rc = l.svsk_ubench_rowshift(win.data_ptr(), rows, r0, mode, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0, l.svsk_last_error()
        torch.cuda.synchronize()
        ref = win[r0:r0 + 128].float()
        ok = torch.equal(out, ref)
        print(f"mode={mode} r0={r0:2d}: {'exact' if ok else 'MISMATCH max|d|=%.3f' % float((out - ref).abs().max())}", flush=True)
