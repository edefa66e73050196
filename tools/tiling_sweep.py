"""Brick height sweep for mid-size boxes (64^3 ... 256^3): which `lz` should `make_brick_grid` pick by default?

One process, CUDA-graph replay as in production (`bflbm_step(h, n)`), fluctuating, tau = 1/2; time with CUDA events.
Usage: python tools/tiling_sweep.py [out.txt]   (GPU only)
"""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bflbm_b200 as B

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)  # torch.cuda.Event records on torch's current stream: make it the lattice's
if len(sys.argv) > 2 and sys.argv[2] == "nt128":  # 128-thread CTAs (experiment knob): small and mid-size boxes only
    os.environ["BFLBM_CTA_THREADS"] = "128"
sizes = [(64, 64, 64), (96, 96, 96), (128, 128, 128), (160, 160, 160), (192, 192, 192), (256, 256, 256), (128, 128, 512), (512, 512, 64)]
kbts = [1e-5, 0.0]
if len(sys.argv) > 2:
    sizes = [(32, 32, 32), (64, 64, 64), (96, 96, 96), (128, 128, 128)] if sys.argv[2] == "nt128" else [s for s in sizes if s[0] in (96, 128, 160, 192)]
    kbts = [1e-5]
for (nx, ny, nz) in sizes:
    cells = nx * ny * nz
    for kbt in kbts:
        best = None
        for lz in [0, 2, 4, 8, 16, 32, 64]:
            if lz > nz:
                continue
            with B.Lattice(nx, ny, nz, B.Params(kBT=kbt, alpha0=1.5, kappa=4.0)) as lat:
                lat.set_stream(stream.cuda_stream)
                lat.set_algorithm("fused")
                if lz:
                    lat.set_tiling(lz)
                lat.init_mixture()
                steps = max(64, min(4096, int(2.0e10 / cells) // 64 * 64))
                lat.step(192)
                lat.sync()
                ts = []
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    lat.step(steps)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) / steps)
                ms = sorted(ts)[1]
                assert lat.check_nan() == 0
            mlups = cells / ms / 1e3
            if lz and (best is None or mlups > best[1]):
                best = (lz, mlups)
            print(f"{nx}x{ny}x{nz} kBT={kbt:g} lz={lz if lz else 'auto'}: {mlups:8.1f} MLUPS  {ms * 1e3:9.2f} us/step", file=out, flush=True)
        print(f"  -> best lz={best[0]} ({best[1]:.1f})", file=out, flush=True)
