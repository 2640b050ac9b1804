"""Bring-up check of the tcgen05 MLP kernels against torch fp32 and the SIMT kernels."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from miccai24_immoco_b200 import _native as nat

torch.backends.cuda.matmul.allow_tf32 = False
lib = mb.lib()


class _Simt:
    """the fp32 SIMT check kernels now live in the test-side library tests/checkers/_mlp_simt.so"""

    def __init__(self):
        import __graft_entry__ as entry
        h = C.CDLL(entry.build_checkers())
        P = C.c_void_p
        h.immoco_simt_mlp_fwd.restype = h.immoco_simt_mlp_bwd.restype = C.c_int
        h.immoco_simt_mlp_fwd.argtypes = [P, P, P, P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, P]
        h.immoco_simt_mlp_bwd.argtypes = [P, P, P, P, P, P, P, C.c_int64, C.c_int32, C.c_int32, P]
        self.immoco_mlp_fwd, self.immoco_mlp_bwd = h.immoco_simt_mlp_fwd, h.immoco_simt_mlp_bwd


_simt = _Simt()


def impl_lib(impl):
    return lib if impl == 1 else _simt


s = lambda: torch.cuda.current_stream().cuda_stream
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())

def run_fwd(enc, w1, w2, width, code, out_tanh, impl):
    n = enc.shape[1]
    out = torch.full((n, 2), float("nan"), device="cuda")
    nat.check(impl_lib(impl).immoco_mlp_fwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), out.data_ptr(), n, width, code, out_tanh, s()), "fwd")
    torch.cuda.synchronize()
    return out

def run_bwd(enc, w1, w2, d_out, width, code, impl):
    n = enc.shape[1]
    d_enc = torch.full_like(enc, float("nan")); g1 = torch.zeros_like(w1); g2 = torch.zeros_like(w2)
    nat.check(impl_lib(impl).immoco_mlp_bwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), d_out.data_ptr(), d_enc.data_ptr(),
                                 g1.data_ptr(), g2.data_ptr(), n, width, code, s()), "bwd")
    torch.cuda.synchronize()
    return d_enc, g1, g2

do_bwd = len(sys.argv) > 1 and sys.argv[1] == "bwd"
ok = True
for width, act in ((64, "tanh"), (256, "relu"), (64, "relu"), (256, "tanh")):
    for n in (128, 1, 1000, 409600 if width == 64 else 102400):
        g = torch.Generator().manual_seed(width + n)
        enc = (torch.randn(16, n, 2, generator=g) * 0.5).cuda()
        w1 = (torch.randn(width, 32, generator=g) * 0.2).cuda().requires_grad_(True)
        w2 = (torch.randn(16, width, generator=g) * 0.2).cuda().requires_grad_(True)
        e = enc.permute(1, 0, 2).reshape(n, 32).clone().requires_grad_(True)
        f = torch.relu if act == "relu" else torch.tanh
        code = nat.ACT_RELU if act == "relu" else nat.ACT_TANH
        ref = (f(e @ w1.t()) @ w2.t())[:, :2]
        for out_tanh in (0, 1):
            r = ref.tanh() if out_tanh else ref
            a = run_fwd(enc, w1, w2, width, code, out_tanh, 1)
            b = run_fwd(enc, w1, w2, width, code, out_tanh, 0)
            ea, eb = rel(a, r.detach()), rel(b, r.detach())
            flag = "OK " if ea < 3e-6 else "BAD"
            ok &= ea < 3e-6
            print(f"{flag} fwd W={width} {act} n={n} out_tanh={out_tanh}: tc rel {ea:.2e}  simt rel {eb:.2e}", flush=True)
        if do_bwd:
            d_out = torch.randn(n, 2, generator=g).cuda()
            (ref * d_out).sum().backward()
            for impl in (1, 0):
                d_enc, g1, g2 = run_bwd(enc, w1, w2, d_out, width, code, impl)
                r1, r2, r3 = rel(d_enc.permute(1, 0, 2).reshape(n, 32), e.grad), rel(g1, w1.grad), rel(g2[:2], w2.grad[:2])
                good = max(r1, r2, r3) < 1e-5 and bool((g2[2:] == 0).all())
                if impl == 1: ok &= good
                print(f"{'OK ' if good else 'BAD'} bwd impl={impl} W={width} {act} n={n}: dE {r1:.2e} gW1 {r2:.2e} gW2 {r3:.2e}", flush=True)
# timing at the C2 sizes
for width, act, n in ((256, nat.ACT_RELU, 102400), (64, nat.ACT_TANH, 409600)):
    enc = torch.randn(16, n, 2, device="cuda") * 0.5
    w1 = torch.randn(width, 32, device="cuda") * 0.2; w2 = torch.randn(16, width, device="cuda") * 0.2
    d_out = torch.randn(n, 2, device="cuda")
    for impl in (0, 1):
        for _ in range(3): run_fwd(enc, w1, w2, width, act, 1, impl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = impl_lib(impl); out = torch.empty((n, 2), device="cuda")
        e0.record()
        for _ in range(20):
            cur.immoco_mlp_fwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), out.data_ptr(), n, width, act, 1, s())
        e1.record(); torch.cuda.synchronize()
        msg = f"time fwd W={width} n={n} impl={impl}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us"
        if do_bwd:
            d_enc = torch.empty_like(enc); g1 = torch.zeros_like(w1); g2 = torch.zeros_like(w2)
            for _ in range(3): run_bwd(enc, w1, w2, d_out, width, act, impl)
            e0.record()
            for _ in range(20):
                cur.immoco_mlp_bwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), d_out.data_ptr(), d_enc.data_ptr(), g1.data_ptr(), g2.data_ptr(), n, width, act, s())
            e1.record(); torch.cuda.synchronize()
            msg += f" | bwd {e0.elapsed_time(e1) / 20 * 1e3:.1f} us"
        print(msg, flush=True)
print("ALL OK" if ok else "FAILURES")
