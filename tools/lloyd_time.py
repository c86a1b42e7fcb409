import sys, time, os
ROOT="/root/repo"
sys.path[:0]=[ROOT, ROOT+"/ann-search-rs_b200/python"]
import numpy as np, annb200
from oracle import datagen
d = datagen.correlated(250_000, 128, seed=42)
init = d[(np.arange(4096)*250_000)//4096].copy()
annb200.kmeans_lloyd(d[:10000], init[:64], annb200.L2, 1)
t=time.perf_counter(); c, it = annb200.kmeans_lloyd(d, init, annb200.L2, 8); dt=time.perf_counter()-t
print(f"device Lloyd 250k x 128, nlist 4096: {it} updates in {dt:.2f} s ({dt/max(it,1)*1e3:.0f} ms per iteration incl. upload)")
