import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import arrowhead_b200 as ab
from helpers import gen_symarrow
ctx = ab.get_context(0)
n = 32768
D, z, a = gen_symarrow(n, seed=421)
k0, k1 = ab.k_partition(n, 8, 3)
U = torch.empty((k1 - k0 + 1, n), dtype=torch.float64, device="cuda")
for rep in range(2):
    G = ctx.symarrow(D, z, a, k0=k0, k1=k1, vectors=False, U_dev=U.data_ptr(), ldu=n)
print(G.stats)
