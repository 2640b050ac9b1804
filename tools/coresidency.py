"""Effect of capping the SM share of the hash-grid / Adam kernels on co-running (one slice and several)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from miccai24_immoco_b200.encoding import grid_spec
from oracle import immoco_oracle as orc
lib = mb.lib(); s = lambda: torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
coords = mb.make_grids((4, 320, 320), "cuda"); gs = grid_spec(3, mb.encoding_config); d = gs.desc(); n = coords.shape[0]
table = (torch.rand(gs.n_rows, 2, device="cuda") - 0.5) * 1e-3
enc = torch.empty(16, n, 2, device="cuda"); d_enc = torch.randn(16, n, 2, device="cuda"); grad = torch.zeros_like(table)
def t(fn, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
n_sl = 4; iters = 200
cases = [orc.make_case(320, 320, 4, 1000 + i) for i in range(n_sl)]
ks = [c["kspace_motion"].to(torch.complex64).pin_memory() for c in cases]
ms = [c["masks"].pin_memory() for c in cases]
mb.reconstruct_batch(ks[:2], ms[:2], 20, in_flight=2); torch.cuda.synchronize()
for hg in (8, 6, 4, 3, 2):
    lib.immoco_set_hashgrid_ctas_per_sm(hg)
    f = t(lambda: lib.immoco_hashgrid_fwd(C.byref(d), coords.data_ptr(), table.data_ptr(), enc.data_ptr(), n, s()))
    b = t(lambda: lib.immoco_hashgrid_bwd(C.byref(d), coords.data_ptr(), d_enc.data_ptr(), grad.data_ptr(), n, s()))
    for av, ac in ((0, 32), (2, 4)):
        lib.immoco_set_adam_tuning(av, ac)
        res = []
        for infl in (1, 2, 4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = mb.reconstruct_batch(ks, ms, iters, in_flight=infl, chunk=10)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            res.append(dt / n_sl / iters * 1e6)
        print(f"hg ctas/SM {hg}: fwd {f:6.1f} bwd {b:6.1f} us | adam(v{av},{ac}/SM) | us per slice-iteration, in_flight 1/2/4: "
              + " / ".join(f"{r:6.1f}" for r in res), flush=True)
