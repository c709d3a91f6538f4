"""Timeline of the tcgen05 attention kernel's CTA 0 (GPU box): per-event clock deltas for the issuer and three
softmax warps.  Needs a library built with PGIBBS_NVCC_EXTRA=-DPGIBBS_FA_TRACE=1 (the
timeline is compiled out of the product kernel).  python tools/attn_trace.py [T]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from protein_gibbs_sampler_b200 import _lib
from protein_gibbs_sampler_b200.engine import op_attention
T = int(sys.argv[1]) if len(sys.argv) > 1 else 258
n_seq, H = 64, 20
qkv = torch.randn(n_seq * T, 3 * H * 64, generator=torch.Generator().manual_seed(1)) * 0.7
lib = _lib.load()
op_attention(qkv, n_seq, T, H, 64)            # warm-up
_lib.check(lib.pgibbs_debug_attention_trace(None, 1))
op_attention(qkv, n_seq, T, H, 64)
buf = np.zeros(4 * 2048, dtype=np.uint64)
_lib.check(lib.pgibbs_debug_attention_trace(buf.ctypes.data_as(ctypes.c_void_p), 0))
names = {0x01: "item", 0x10: "seenP_A", 0x11: "seenP_B", 0x18: "issued_A", 0x19: "issued_B", 0x20: "waitS", 0x21: "gotS",
         0x22: "gaveP", 0x30: "fenced", 0x31: "pv_issued", 0x32: "pv_committed", 0x23: "O_full", 0x24: "O_staged", 0x25: "grp_bar"}
t0 = min(int(v >> np.uint64(8)) for v in buf if v)
for slot, who in enumerate(["issuer", "softmax w4 (A, quad0)", "softmax w5 (A, quad1)", "softmax w8 (B, quad0)"]):
    ev = [(int(v >> np.uint64(8)) - t0, int(v & np.uint64(0xff))) for v in buf[slot * 2048:(slot + 1) * 2048] if v]
    print("==", who, len(ev), "events")
    prev = None
    for k, (t, c) in enumerate(ev[:int(os.environ.get("N_EV", 140))]):
        print("%8d  +%6d  %s" % (t, 0 if prev is None else t - prev, names.get(c, hex(c))))
        prev = t
