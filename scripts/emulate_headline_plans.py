"""Replay the pushed halo exchange of the persistent kernel on the CPU for the plans the headline benchmark uses
(1920x2520 RGB, Gaussian, 100 iterations; 1, 2, 4, 8 ranks; 8 and 16 warps per CTA): every rank, CTA, warp and lane
goes through the device code's own index functions (csrc/common/xplan.h via csrc/host/xplan_emu.cpp) and the
result must equal the oracle byte for byte. CPU only; about a minute per plan.

    python scripts/emulate_headline_plans.py [loops]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np, sys
from parallel_convolution_b200.models.filters import Filter
from parallel_convolution_b200.ops import host as H
from parallel_convolution_b200.parallel.plan import plan_persist
f=Filter.builtin("gaussian")
Hh,W,C=2520,1920,3
img=np.random.default_rng(0).integers(0,256,(Hh,W,C),dtype=np.uint8)
loops=int(sys.argv[1]) if len(sys.argv)>1 else 100
want=H.oracle(img,f.spec,loops)
for n,grid in ((1,(1,1)),(2,(2,1)),(4,(2,2)),(8,(4,2)),(8,(8,1))):
    tr,tc=Hh//grid[0],W//grid[1]
    for warps in (16,8):
        p=plan_persist(tr,tc*C,C,f.spec.kind,loops,min_nbr_rows=tr if grid[0]>1 else None,min_nbr_cols=tc if grid[1]>1 else None,peers=n>1,warps=warps)
        if p is None: print(n,grid,warps,"no plan"); continue
        t=time.time()
        got=H.xplan_emulate(img,f.spec,loops,grid,p.k,p.R,warps,p.TH,3)
        ok = (not isinstance(got,int)) and np.array_equal(got,want)
        print(f"N={n} grid={grid} W={warps} plan k={p.k} R={p.R} TH={p.TH} TWB={p.TWB} nty x ntx={p.nty}x{p.ntx}: {'== oracle' if ok else 'MISMATCH/'+str(got if isinstance(got,int) else '')}  ({time.time()-t:.1f}s)", flush=True)
