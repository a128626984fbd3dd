import sys, ctypes as C, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sgmcmc_b200 import _lib
import sgmcmc_b200 as S
w=bench.workload(sys.argv[1] if len(sys.argv)>1 else 'c3', None)
sg=bench.make_sampler(w,0,1,w['chains'])
lib=_lib.load()
lib.sgm_debug_phase_cycles.argtypes=[C.c_void_p,C.c_int,C.POINTER(C.c_ulonglong),C.c_int]
S.initSess(sg,verbose=False)
_lib.check(lib.sgm_run(sg.handle,50 if w['family']!='nn' else 2,0))
buf=(C.c_ulonglong*16)()
lib.sgm_debug_phase_cycles(sg.handle,1,None,0)
NIT=100 if w['family']!='nn' else 5
_lib.check(lib.sgm_run(sg.handle,NIT,0))
lib.sgm_debug_phase_cycles(sg.handle,1,buf,16)
v=np.array(list(buf),dtype=np.float64)
names=['store/loop','sync pre-gather','idx+issue','wait rows','phase1','sigmoid/coef','block sum','phase2','update','sync post']
if w['family']=='nn': names=['S1 (total)','S2','S3 gradA','S4','S5 gradB GEMM','update A,a,b','eval prologue','store','  (S1: wait loads)','  (S1: wait lo stage)','update B','update lambdas']
steps=w['chains']*NIT
for n,x in zip(names,v): print(f"{n:18s} {x/steps:10.0f} cycles/chain-step {100*x/v.sum():5.1f}%")
print('total (S1 sub-counters excluded)', (v.sum()-v[8]-v[9])/steps)
