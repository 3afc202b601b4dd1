import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from studiosr_b200.models import SwinIR
from oracle.synth import smooth_image_u8
torch.manual_seed(0)
m = SwinIR(scale=4).cuda().eval(); m.precision = "bf16"
img = smooth_image_u8(1080, 1920, seed=5)
out = m.inference_tiled(img)          # warm-up + correctness of shapes
torch.cuda.synchronize(); t = time.perf_counter()
out = m.inference_tiled(img)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("1920x1080 -> ", out.shape, out.dtype, "finite", np.isfinite(out.astype(np.float32)).all(), f"{dt*1e3:.1f} ms", f"{out.shape[0]*out.shape[1]/1e6/dt:.1f} Mpix/s e2e")
# consistency: the top-left 64x64 tile region far from seams equals the single-tile path? compare a crop with a smaller frame
small = m.inference_tiled(np.ascontiguousarray(img[:540, :960]))
d = np.abs(out[:1900, :3600].astype(int) - small[:1900, :3600].astype(int))
print("max |diff| vs 960x540 sub-frame on the shared interior:", d[:1700, :3400].max())
