"""GPU probe: time of one warm-started split through the C ABI (CUDA events, alone on the GPU) and the in-kernel phase
clocks of k_fast_split; bond-like matrices as in tests/test_gpu_fast_split.py."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from tensornetworkforml_b200 import _lib as L
from test_gpu_fast_split import bond_like

L.lib()
Dl = Dr = m = int(os.environ.get("D", 64)); nl = int(os.environ.get("NL", 10))
NSV = 2 * Dl
for left_dir in (0, 1):
    for spread in (0.5, 0.997):
        rng = np.random.default_rng(1)
        lib = L.lib()
        ws = torch.empty(lib.tnml_svd_split_workspace_bytes(Dl, Dr, nl, left_dir) // 8 + 1, dtype=torch.float64, device="cuda")
        warm = torch.zeros(lib.tnml_svd_warm_bytes(Dl, Dr, nl, left_dir) // 8, dtype=torch.float64, device="cuda")
        site_p = torch.empty(Dl * 2 * m * nl, dtype=torch.float64, device="cuda")
        site_q = torch.empty(m * 2 * Dr * nl, dtype=torch.float64, device="cuda")
        sv = torch.zeros(4096, dtype=torch.float64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        Mx = bond_like(rng, Dl, Dr, nl, left_dir, spread)
        for visit in range(4):
            Bd = torch.from_numpy(np.ascontiguousarray(Mx)).to("cuda")
            fast = 1 if visit > 0 else 0
            ts = []
            for rep in range(3):
                wsave = warm.clone()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                L.call("tnml_svd_split_warm", Bd.data_ptr(), site_p.data_ptr(), site_q.data_ptr(), sv.data_ptr(), ws.data_ptr(),
                       warm.data_ptr(), Dl, Dr, nl, m, left_dir, 3, fast, L.F64, st, None)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
                if rep < 2:
                    warm.copy_(wsave)
            s = sv.cpu().numpy()
            print("left_dir=%d spread=%.3f visit %d fast=%d: %s us, marker %d, phase cycles %s" % (
                left_dir, spread, visit, fast, ["%.0f" % t for t in ts], int(s[NSV]), [int(x) for x in s[NSV + 4:NSV + 14]]), flush=True)
            Mx = Mx + 1e-3 * bond_like(rng, Dl, Dr, nl, left_dir, spread)
