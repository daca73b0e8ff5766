import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import numpy as np, scipy.sparse as sp
from hmclab_b200 import workloads
w = workloads.tomography()
lik = [d for d in w.posterior.separate_distributions if type(d).__name__=="LinearMatrix"][0].Distribution
G = sp.csr_matrix(lik.G)
print(G.shape, G.nnz)
def stats(G, R=128):
    Gb = sp.csr_matrix((np.ones_like(G.data), G.indices, G.indptr), shape=G.shape)
    GbT = Gb.T.tocsr()
    rows = G.shape[0]
    assigned = np.zeros(rows, bool)
    Ks = []; sizes=[]
    t0=time.time()
    for seed in range(rows):
        if assigned[seed]: continue
        assigned[seed] = True
        cnt = np.asarray((Gb[seed] @ GbT).todense()).ravel()
        cnt[assigned] = 0
        cand = np.argpartition(-cnt, R-1)[:R-1]
        cand = cand[cnt[cand] > 0]
        assigned[cand] = True
        members = np.concatenate([[seed], cand])
        cols = np.unique(G[members].indices)
        Ks.append(len(cols)); sizes.append(len(members))
    Ks=np.array(Ks); sizes=np.array(sizes)
    kpad = ((Ks+127)//128)*128
    print("bundles", len(Ks), "full", (sizes==R).sum(), "rows in partial", sizes[sizes<R].sum(), "mean K", Ks.mean(), "sum Kpad", kpad.sum(), "ideal bundles", rows/R, "time", time.time()-t0)
    print("K percentiles", np.percentile(Ks,[10,50,90,99]), "size pct", np.percentile(sizes,[10,50,90]))
    return Ks, sizes
stats(G)
stats(sp.csr_matrix(G.T))
