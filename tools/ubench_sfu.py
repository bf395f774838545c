"""Special-function-unit throughput per SM (tanh / ex2 / rcp) and of candidate formulations of the gate
z = tanh(a) * sigmoid(b), in results per clock per SM — what bounds the gating epilogues of the block kernels."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ensemble_svs_with_interactions_b200.csrc.build import build_ubench  # noqa: E402

l = C.CDLL(build_ubench())
l.svsk_ubench_sfu.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
l.svsk_last_error.restype = C.c_char_p
names = {0: "tanh.approx.f32", 1: "ex2.approx + FADD", 2: "rcp.approx + FADD", 3: "FFMA", 4: "gate: 2 x MUFU.TANH (kernels today)",
         5: "gate: 2 x EX2 + RCP", 6: "gate: 2 x exp2 on the FMA pipe + RCP", 7: "gate: EX2 + FMA-pipe exp2 + RCP"}
iters, grid = 2000, 148
sink = torch.zeros(4, device="cuda")
for mode, name in names.items():
    for warps in (4, 8, 16):
        out = torch.zeros(grid, dtype=torch.int64, device="cuda")
        for _ in range(2):
            rc = l.svsk_ubench_sfu(mode, warps, iters, grid, out.data_ptr(), sink.data_ptr(), torch.cuda.current_stream().cuda_stream)
            assert rc == 0, l.svsk_last_error()
        torch.cuda.synchronize()
        cyc = out.float().mean().item()
        n = warps * 32 * 8 * iters
        print(f"{name:42s} {warps:2d} warps: {n / cyc:6.2f} results/clk/SM  ({cyc / iters / 8:6.2f} cycles per warp-wide result per warp)", flush=True)

# TMEM read port: bytes per clock per SM
l.svsk_ubench_tmem_ld.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
for width in (16, 32):
    for piw in (0, 1):
        for warps in (4, 8, 16):
            out = torch.zeros(grid, dtype=torch.int64, device="cuda")
            for _ in range(2):
                rc = l.svsk_ubench_tmem_ld(width, piw, warps, 500, grid, out.data_ptr(), sink.data_ptr(), torch.cuda.current_stream().cuda_stream)
                assert rc == 0, l.svsk_last_error()
            torch.cuda.synchronize()
            cyc = out.float().mean().item()
            print(f"tcgen05.ld 32x32b.x{width} wait per {'instruction' if piw else 'round of 128 columns'} {warps:2d} warps: "
                  f"{warps * 16384 * 500 / cyc:6.1f} B/clk/SM", flush=True)
