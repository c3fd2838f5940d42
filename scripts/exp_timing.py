"""Role timers of conv_tc_kernel (A/B build -DIEVM_EXP_TIMING, ab/lib_timing*.so via IEVM_LIB_PATH): per conv layer, mean over
CTAs of the cycles the MMA warp / TMA producer / one epilogue warp spend in their loops and in their mbarrier waits, the
whole-kernel cycles, and the SM clock the kernel actually ran at (cycles / globaltimer ns).
usage: IEVM_LIB_PATH=ab/lib_timing.so python scripts/exp_timing.py [N]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ievm_b200
from ievm_b200 import synthetic as mf, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dtype = sys.argv[2] if len(sys.argv) > 2 else "i8"
if dtype == "i8":
    eng = ievm_b200.B200QuantizedResNet.from_converted(mf.static_quantize_fbgemm(mf.make_student()), max_batch=n)
    x = mf.synthetic_images(n).cuda()
else:
    eng = ievm_b200.B200HalfResNet.from_half_module(mf.cast_fp16(mf.make_student()), max_batch=n)
    x = mf.synthetic_images(n).half().cuda()
lib = _lib.load()
SLOTS, CTAS, WORDS = 32, 160, 16
buf = np.zeros(SLOTS * CTAS * WORDS, dtype=np.uint64)
for _ in range(3):
    eng(x)
lib.ievm_exp_timing_read(C.c_void_p(buf.ctypes.data), 1)
eng(x)
lib.ievm_exp_timing_read(C.c_void_p(buf.ctypes.data), 1)
t = buf.reshape(SLOTS, CTAS, WORDS).astype(np.float64)
print("layer                 ctas tiles/cta | kernel clk   MHz | mma: loop  w.tempty  w.full | tma: loop  w.empty | epi: loop  w.tfull")
for li, L in enumerate(eng.net.layers):
    a = t[li % SLOTS]
    live = a[:, 9] > 0
    if not live.any():
        continue
    m = a[live].mean(0)
    mhz = 1e3 * a[live][:, 9].sum() / max(a[live][:, 10].sum(), 1)
    print(f"{L.name:22s} {int(live.sum()):4d} {4 * m[8]:8.1f}  | {m[9]:9.0f} {mhz:6.0f} | {m[0]:9.0f} {m[1]:9.0f} {m[2]:8.0f} | {m[4]:9.0f} {m[5]:8.0f} | "
          f"{m[6]:9.0f} {m[7]:8.0f}")

a = t[31]
live = a[:, 10] > 0
if live.any():
    m = a[live].mean(0)
    print(f"front end (frontend2_kernel), {int(live.sum())} CTAs, {m[8]:.1f} tiles/CTA:")
    print(f"  MMA warp   loop {m[0]:9.0f}  wait (both barriers) {m[2]:8.0f}  of which blocking wait for records {m[1]:8.0f}")
    print(f"  TMA warp   loop {m[4]:9.0f}  wait raw slot free {m[5]:8.0f}")
    print(f"  quantiser  loop {m[11]:9.0f}  wait raw rows {m[12]:8.0f}  wait line slot free {m[13]:8.0f}")
    print(f"  epilogue   loop {m[6]:9.0f}  wait accumulator {m[7]:8.0f}")
