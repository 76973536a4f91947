import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import snprelate_b200 as S
from oracle import snprel_oracle as O
import importlib.util
spec = importlib.util.spec_from_file_location("tgp", "/root/repo/tests/test_gpu_parity.py")
ctx = S.Context(0)
src = open('/root/repo/tests/test_gpu_parity.py').read()
ns = {}
exec("import numpy as np\n" + "def _two_populations" + src.split("def _two_populations")[1].split("@pytest.mark.parametrize")[0], ns)
for structured in (True, False):
    t0=time.time()
    g = ns['_two_populations'](2304, 6000, 11) if structured else O.synth_geno(2304, 6000, seed=21, miss_rate=0.002)
    t1=time.time()
    ctx.geno_begin(g.shape[1], g.shape[0]); ctx.geno_push_u8(g)
    t2=time.time()
    r1 = ctx.pca(eigen_cnt=16, need_genmat=True)
    t3=time.time()
    print(structured, 'gen %.1fs push %.1fs pca %.2fs'%(t1-t0,t2-t1,t3-t2), ctx.last_eigen_info(), ctx.eigen_phase_ms, flush=True)
    ctx.debug_flags(4); t4=time.time(); r0=ctx.pca(eigen_cnt=16); t5=time.time(); ctx.debug_flags(0)
    print('  dense pca %.2fs'%(t5-t4), flush=True)
    t6=time.time(); val, vec = O.pca_eigen(r1["genmat"], 16); print('  oracle eigh %.1fs'%(time.time()-t6), flush=True)
