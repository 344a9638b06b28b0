#!/usr/bin/env python
"""Timeline of CTA 0 of attn_tcp_kernel (library built with -DMESM_ATP_TRACE: make -C mesm_b200/csrc clean all EXTRA=-DMESM_ATP_TRACE).

Events per head: loader 0 K/Q tiles free, 1 K/Q staged, 2 V free, 3 V staged; MMA thread 4 K/Q seen, 5 S MMAs issued, 6 first P block
seen, 7 V seen, 8 last P V issued; softmax warp 1: 9 S ready, 10 row max done, 11 last P block written, 12 O ready, 13 output stored.
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesm_b200 import _lib  # noqa: E402

B, L = 296, 147
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * L, 768, device="cuda", generator=g)
pad = torch.zeros(B, L, dtype=torch.bool, device="cuda")
pad[:, 0] = True
out = torch.empty(B * L, 256, device="cuda")
lib = _lib.lib()
wd = (ctypes.c_ulonglong * 8)()
p = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(3):
    lib.mesm_debug_attention(p(qkv), p(pad.view(torch.uint8)), B, L, p(out), 1, 1, wd, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
tr = (ctypes.c_longlong * 128)()
n = lib.mesm_debug_attn_trace(tr)
if not n:
    sys.exit("library was not built with -DMESM_ATP_TRACE")
t0 = min(v for v in tr if v > 0)
names = ["ld:KQfree", "ld:KQdone", "ld:Vfree", "ld:Vdone", "mma:KQseen", "mma:Sissued", "mma:P0seen", "mma:Vseen", "mma:PVdone", "sm:Sready",
         "sm:max", "sm:Pdone", "sm:Oready", "sm:stored"]
print("head " + " ".join(f"{n:>11s}" for n in names))
for h in range(8):
    print(f"{h:4d} " + " ".join(f"{tr[h * 16 + e] - t0:11d}" for e in range(14)))
