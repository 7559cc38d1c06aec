import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle as O
from wise_b200 import faiss_compat as faiss
from wise_b200.sharded import train_ivf_sharded
n, d, k = 20000, 64, 100
x = O.clustered_unit(n, d, 100, 31)
ivf = faiss.IndexIVFFlat(faiss.IndexFlatIP(d), d, k, faiss.METRIC_INNER_PRODUCT)
objs = train_ivf_sharded(ivf, torch.from_numpy(x).cuda(), verbose=True)
print(objs)
