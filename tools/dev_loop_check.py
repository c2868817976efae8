import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from common import *
from swiftlink_b200 import capi
import test_gpu_parity as T
fx = golden('loop'); plan = capi.Plan(problem('loop'))
ref = np.load('/root/repo/tests/golden/loop_ref_lsampler_300.npy')
mine = np.stack([T._gpu_chain_lod(plan, fx, 5000+s) for s in range(120)])
a, b = mine.mean(axis=(1,2)), ref.mean(axis=(1,2))
print('device mean %.4f sd %.4f se %.4f | ref mean %.4f sd %.4f se %.4f | z %.2f' % (a.mean(), a.std(ddof=1), a.std(ddof=1)/np.sqrt(len(a)), b.mean(), b.std(ddof=1), b.std(ddof=1)/np.sqrt(len(b)), abs(a.mean()-b.mean())/np.sqrt(a.var(ddof=1)/len(a)+b.var(ddof=1)/len(b))))
print('per position device', mine.mean(0).round(3).ravel()); print('per position ref   ', ref.mean(0).round(3).ravel())
