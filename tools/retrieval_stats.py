#!/usr/bin/env python
"""Where the epilogue warps of retrieve_topk_kernel spend their cycles (diagnostic knob retrieval_diag = 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from manner_b200 import ops, retrieval  # noqa: F401  (retrieval registers torch.ops.manner_b200.retrieve_topk)

dev = torch.device("cuda:0")
n_users, n_cat, dim = int(sys.argv[1]) if len(sys.argv) > 1 else 37888, int(sys.argv[2]) if len(sys.argv) > 2 else 1_250_000, 768
g = torch.Generator(device=dev).manual_seed(1)
cat = (torch.randn(n_cat, dim, generator=g, device=dev) * dim ** -0.5).to(torch.bfloat16)
users = (torch.randn(n_users, dim, generator=g, device=dev) * dim ** -0.5).to(torch.bfloat16)
ops.set_tuning(retrieval_diag=4)
for it in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.ops.manner_b200.retrieve_topk(users, cat, 100, 0, False); e1.record(); torch.cuda.synchronize()
    ws = [v for k, v in ops._workspaces.items() if k[2] == "retrieval"][0]
    st = ws[64:64 + 64].view(torch.int64).cpu().tolist()
    warps = 148 * 8
    print(f"ms {e0.elapsed_time(e1):.2f}  per epilogue warp (Mcycles): total {st[5]/warps/1e6:.2f} compaction {st[0]/warps/1e6:.2f} slow path {st[1]/warps/1e6:.2f} "
          f"wait tmem_full {st[2]/warps/1e6:.2f} | compactions/warp {st[3]/warps:.0f} slow chunks/warp {st[4]/warps:.0f} | MMA thread per CTA: "
          f"waits for a free accumulator {st[6]/148/1e6:.2f}, for operands {st[7]/148/1e6:.2f}")
