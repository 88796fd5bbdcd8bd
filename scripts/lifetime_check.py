import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from multiviewstitch_b200 import api
ctx = api.Context(0)
idx = api.NNIndex(ctx).build(np.random.rand(1000,3).astype(np.float32))
s = api.SRTSolver(10, ctx)
ctx.close()          # context first
print(idx.knn(np.random.rand(10,3).astype(np.float32), 1)[0].shape)
idx.close(); s.close()
ctx2 = api.Context(0)
idx2 = api.NNIndex(ctx2).build(np.random.rand(1000,3).astype(np.float32))
print("exit without closing anything")
