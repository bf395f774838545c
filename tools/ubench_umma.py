"""Cycles per tcgen05.mma (SS mode, bf16, K = 16) with operands resident in smem — ground truth for the kernels' pipe time."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ensemble_svs_with_interactions_b200.csrc.build import build_ubench  # noqa: E402

l = C.CDLL(build_ubench())   # tools/libsvsk_ubench.so — the micro-benchmarks are not part of the product libsvsk.so
l.svsk_ubench_umma.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
l.svsk_last_error.restype = C.c_char_p
iters = 2000
for grid in (2, 148):
    for cg, N in ((1, 64), (1, 96), (1, 128), (1, 256), (2, 64), (2, 128), (2, 256)):
        for adv in (0, 1):
            out = torch.zeros(grid * 2, dtype=torch.int64, device="cuda")
            rc = l.svsk_ubench_umma(cg, N, iters, adv, grid, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            assert rc == 0, l.svsk_last_error()
            torch.cuda.synchronize()
            o = out.view(grid, 2).float()
            o = o[o[:, 1] > 0]
            M = 128 * cg
            nominal = 128 * N / 256  # cycles per MMA per SM at 4096 MAC/cycle/SM
            print(f"grid={grid:3d} cta_group={cg} M={M} N={N:3d} advance={adv}: issue {o[:, 0].mean() / iters:6.1f} "
                  f"cyc/MMA, complete {o[:, 1].mean() / iters:6.1f} cyc/MMA (nominal {nominal:.0f}); "
                  f"operand bytes/MMA/SM = {(128 + (N if cg == 1 else N // 2)) * 32}", flush=True)

# commit latency: groups of G MMAs + commit + wait, serially
for cg, N in ((1, 128), (2, 256)):
    for G in (1, 4, 8):
        out = torch.zeros(4, dtype=torch.int64, device="cuda")
        it = 500
        rc = l.svsk_ubench_umma(cg, N, it, 16 + G - 1, 2, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0, l.svsk_last_error()
        torch.cuda.synchronize()
        per = float(out[0]) / it
        nominal = 128 * cg * N / (256 * cg) * G
        print(f"serial cta_group={cg} N={N} group={G}: {per:7.1f} cycles per (group + commit + wait); MMAs nominal {nominal:.0f} "
              f"-> commit-to-visible latency ~ {per - nominal:.0f} cycles", flush=True)

# issue loop with per-group extras (see ubench_sm100.cu): does anything in the MMA thread's loop stall behind the MMAs?
names = {0: "MMAs only", 4: "+commit", 1: "+2 try_wait", 9: "+2 test_wait", 3: "+2 try_wait +fence", 7: "+2 try_wait +fence +commit",
         15: "+2 test_wait +fence +commit", 6: "+fence +commit"}
for cg, N in ((1, 128), (2, 256)):
    for v, nm in names.items():
        out = torch.zeros(4, dtype=torch.int64, device="cuda")
        it = 500
        rc = l.svsk_ubench_umma(cg, N, it, 32 + v, 2, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0, l.svsk_last_error()
        torch.cuda.synchronize()
        print(f"issue loop cta_group={cg} N={N} [{nm:32s}]: issue {float(out[0]) / it:7.1f}  complete {float(out[1]) / it:7.1f} cycles per "
              f"group of 4 MMAs (nominal {4 * 128 * N // 256})", flush=True)
