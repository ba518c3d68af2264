"""cuSOLVER potrf variants at the S5 size (diagnostic): library entry point (cusolverDnDpotrf) vs torch.linalg.cholesky_ex."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpr_calculator_b200 import _lib                      # noqa: E402
from gpr_calculator_b200.device import ptr, stream        # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32980
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn((N, 2048), dtype=torch.float64, device="cuda", generator=g)
K0 = A @ A.T
K0 += N * torch.eye(N, dtype=torch.float64, device="cuda")
del A
names = ("gprb_chol_factor", "torch.linalg.cholesky_ex(upper=False)", "torch.linalg.cholesky_ex(upper=True)")
if len(sys.argv) > 2 and sys.argv[2] == "lib":
    names = names[:1]
for name in names:
    ts = []
    for it in range(3):
        K = K0.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if name == "gprb_chol_factor":
            _lib.call("gprb_chol_factor", ptr(K), N, N, stream())
        else:
            torch.linalg.cholesky_ex(K, upper=name.endswith("True)"), out=(K, torch.empty((), dtype=torch.int32, device="cuda")))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        del K
    print("%-42s N=%d  %s ms  (%.1f TFLOP/s)" % (name, N, ["%.1f" % t for t in ts], N ** 3 / 3 / (min(ts) * 1e-3) * 1e-12), flush=True)
