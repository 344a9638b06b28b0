"""Time the self-attention kernels on one bench-sized launch: python tools/attn_bench.py [pairs] [L]."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesm_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = int(sys.argv[2]) if len(sys.argv) > 2 else 147
lib = _lib.lib()
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * L, 768, device="cuda", generator=g)
pad = torch.zeros(B, L, dtype=torch.uint8, device="cuda"); pad[:, 0] = 1
out = torch.empty(B * L, 256, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr())
wd = (ctypes.c_ulonglong * 8)()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for mode, name in ((3, "attn_mma"), (1, "attn_tcp"), (0, "mha_rows")):
    if mode == 1 and (L > 224 or L <= 64): continue
    lib.mesm_debug_attention(p(qkv), p(pad), B, L, p(out), mode, 2, wd, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 10
    e0.record(); rc = lib.mesm_debug_attention(p(qkv), p(pad), B, L, p(out), mode, it, wd, st); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    print(f"{name}: B={B} L={L} rc={rc} {ms*1e3:.1f} us/launch  {B*L*4096/ms/1e6:.0f} GB/s algorithmic")
