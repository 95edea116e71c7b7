"""Development helper: is the system-reserved first KB of shared memory (where the tcgen05.alloc/dealloc
sequences keep their bookkeeping) re-initialised for every CTA?  Pokes it from a dummy kernel on every SM and
runs a tensor-core batch search afterwards."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch, cqs_b200
import bench as B
from cqs_b200.capi import lib
n = 1_000_000
dev = torch.device("cuda", 0)
ix = cqs_b200.B200Index(768, storage="bf16")
ix.reserve(n)
for b in range(n // B.BLK):
    x = B.gen_block(torch, dev, b, "uniform")
    ix.append_device(x.data_ptr(), x.shape[0])
ix.finalize()
q = B.make_queries(256, 11)
lib.cqs_b200_debug_poke_reserved.argtypes = [C.c_uint32] * 5
ix.search_batch_rows(q, 100)
print("batch search before the poke: ok", flush=True)
for lo, hi, dyn in ((0x40, 0x60, 0), (0x40, 0x60, 200 * 1024), (0x0, 0x400, 200 * 1024)):
    rc = lib.cqs_b200_debug_poke_reserved(148 * 8, 0xFFFFFFFF, lo, hi, dyn)
    print(f"poke [{lo:#x}, {hi:#x}) dyn {dyn}: rc {rc}", flush=True)
    try:
        ix.search_batch_rows(q, 100)
        print("  batch search after it: ok", flush=True)
    except Exception as e:
        print("  batch search after it:", str(e)[-120:], flush=True)
        break
