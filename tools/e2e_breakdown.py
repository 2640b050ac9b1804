"""Where does the public-API call spend its time? (host wall clock, synchronised phases)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc

def t():
    torch.cuda.synchronize(); return time.perf_counter()

case = orc.make_case(320, 320, 4, 1004)
k = case["kspace_motion"].pin_memory(); masks = case["masks"].pin_memory()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
for rep in range(3):
    t0 = t()
    m_dev = masks.cuda(); t1 = t()
    model = mb.IMMoCo(m_dev); t2 = t()
    kc = k.cuda(); scale = kc.abs().max(); kin = kc.div(scale).mul(16000).clone(); t3 = t()
    lam = mb.lambda_schedule(iters, 1e-2)
    eng = mb.FitEngine(model, iters); eng.set_kspace(kin); t4 = t()
    h0 = time.perf_counter(); eng.run(lam, 1e-2); h1 = time.perf_counter(); t5 = t()
    img = torch.view_as_complex(eng.image.clone()).cpu(); t6 = t()
    del eng, model; torch.cuda.empty_cache(); t7 = t()
    print(f"rep {rep}: masks h2d {1e3*(t1-t0):.1f} ms | IMMoCo() {1e3*(t2-t1):.1f} | k norm {1e3*(t3-t2):.1f} | "
          f"FitEngine() {1e3*(t4-t3):.1f} | run {1e3*(t5-t4):.1f} (host enqueue {1e3*(h1-h0):.1f}) | "
          f"d2h {1e3*(t6-t5):.1f} | free {1e3*(t7-t6):.1f} | total {1e3*(t7-t0):.1f}")
t0 = t(); im, kf = mb.imcoco_motion_correction(k, masks, iters); im = im.cpu(); t1 = t()
print(f"api call total {1e3*(t1-t0):.1f} ms")
