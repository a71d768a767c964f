"""Development probe: ulp error of the straight-line elementary functions of a solver library."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tpl_b200 import _cabi
lib = _cabi.load(sys.argv[1])
rng = np.random.default_rng(0)
n = 1 << 20
angles = np.concatenate([rng.uniform(-10.0, 10.0, n // 2), rng.uniform(-1e5, 1e5, n // 4), rng.uniform(-1e-3, 1e-3, n // 4)])
positive = np.concatenate([rng.uniform(1e-12, 1.0, n // 2), rng.uniform(1.0, 1e12, n // 2)])
signed = positive * rng.choice([-1.0, 1.0], n)
def ulp(a, b):
    return float(np.max(np.abs(a - b) / np.spacing(np.abs(b))))
for fn, xs, ref, name in [(0, angles, np.sin, "sin"), (1, angles, np.cos, "cos"), (2, angles, np.tan, "tan"),
                          (3, signed, lambda v: 1.0 / v, "inv"), (4, positive, lambda v: 1.0 / np.sqrt(v), "rsqrt"), (5, positive, np.sqrt, "sqrt")]:
    x = torch.from_numpy(xs).cuda(); out = torch.empty_like(x)
    _cabi.check(lib, lib.tplb_selftest_math(fn, x.data_ptr(), x.numel(), out.data_ptr(), torch.cuda.current_stream().cuda_stream), "selftest")
    print(f"{name}: {ulp(out.cpu().numpy(), ref(xs)):.2f} ulp")
