#!/usr/bin/env python
"""Fitch parsimony on wide trees: one full-traversal pll_fastparsimony_update_vectors (wall time of the blocking
call) and, with a third argument, the stepwise-addition tree.

  python profiles/tools/bench_parsimony_wide.py TIPS SITES [s]
  PLF_PARS_LEVELS=0 ...   always the one-launch chain kernel (the list is as deep as it is long)
  default                 level-scheduled when the list is several times longer than deep (profiles/r1_notes.md)
"""
import sys, importlib, time, json, ctypes as C, os
sys.path.insert(0,'/root/repo')
import numpy as np
pkg=importlib.import_module('libpll-2_b200'); capi=pkg.capi
harness=importlib.import_module('libpll-2_b200.harness'); synth=importlib.import_module('libpll-2_b200.synth')
tips,sites=int(sys.argv[1]),int(sys.argv[2])
ds=synth.dna_dataset(tips,sites,seed=3,alpha=0.3,brlen=(0.002,0.05),simulate_down_tree=False)
lib=pkg.load()
eng=harness.Engine(lib,ds,capi.ARCH_CUDA|capi.PATTERN_TIP)
p=lib.pll_fastparsimony_init(eng.p)
triples=[(int(r[0]),int(r[2]),int(r[5])) for r in ds.tree.ops]
ops=(capi.ParsBuildOp*len(triples))(*[capi.ParsBuildOp(*t) for t in triples])
for _ in range(3): lib.pll_fastparsimony_update_vectors(p,ops,len(triples))
t0=time.perf_counter()
for _ in range(10): lib.pll_fastparsimony_update_vectors(p,ops,len(triples))
dt=(time.perf_counter()-t0)/10*1e3
a,b=ds.tree.root_edge[:2]
out={"tips":tips,"sites":sites,"levels":os.environ.get("PLF_PARS_LEVELS","auto"),"traversal_ms":dt,"score":lib.pll_fastparsimony_edge_score(p,a,b)}
if len(sys.argv)>3:
    labels=(C.c_char_p*tips)(*[f"t{i}".encode() for i in range(tips)])
    arr=(capi.ParsimonyP*1)(p); cost=C.c_uint(0)
    t0=time.perf_counter(); tree=lib.pll_fastparsimony_stepwise(arr,labels,C.byref(cost),1,1)
    out["stepwise_ms"]=(time.perf_counter()-t0)*1e3; out["cost"]=cost.value
print(json.dumps(out))
