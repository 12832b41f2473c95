"""Time one pir_pwdw launch (diagnostic).  python tools/time_pwdw.py B H W C N gate [reps]
Honours PIR_PWDW_CFG / PIR_PWDW_DBG (read by the library at launch time)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promptir_b200 import ops, packing  # noqa: E402
from promptir_b200._lib import LN_WITHBIAS  # noqa: E402


def main():
    B, H, W, C, N, gate = (int(v) for v in sys.argv[1:7])
    reps = int(sys.argv[7]) if len(sys.argv) > 7 else 10
    dt = torch.float16 if os.environ.get('PIR_TIME_DTYPE') == 'fp16' else torch.bfloat16
    torch.manual_seed(0)
    npre = 2 * N if gate else N
    x = torch.randn(B, H, W, C, device="cuda").to(dt)
    w16, _, vec_t = packing.pack_pointwise(torch.randn(npre, C, device="cuda") / C ** 0.5, dt, gamma=torch.ones(C, device="cuda"),
                                           beta=torch.zeros(C, device="cuda"))
    dw = packing.pack_depthwise(torch.randn(npre, 1, 3, 3, device="cuda") / 3, torch.float16)
    out = torch.zeros(B, H, W, N, device="cuda", dtype=dt)
    launch = ops.pwdw(x, w16, dw, out, gate=bool(gate), ln_mode=LN_WITHBIAS, vec_t=vec_t)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        launch(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch(s)
    e1.record()
    torch.cuda.synchronize()
    print(f"pwdw B{B} {H}x{W} C{C} N{N} gate{gate} cfg={os.environ.get('PIR_PWDW_CFG', '-')} dbg={os.environ.get('PIR_PWDW_DBG', '0')}: "
          f"{e0.elapsed_time(e1) / reps * 1e3:.1f} us")


if __name__ == "__main__":
    main()
