"""Quick GPU check of the selection kernels with long candidate lists (capacity > 4096: dynamic shared memory above the
default limit): k = 100 with a 1/64 threshold sample on a 49-tile shard, pre-filter on vs off (the f64 path)."""
import os, sys
os.environ["B2R_SAMPLE_STEP"] = "64"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import b200ret
from b200ret import synthetic as S
n_docs, n_vocab, k = 200_000, 20_000, 100
data, indices, indptr, dl = S.zipf_corpus(n_docs, n_vocab, 60, seed=5)
idf = b200ret.reference_idf(indices, n_docs, n_vocab); avgdl = b200ret.reference_avgdl(dl)
q = S.zipf_queries(128, n_vocab, seed=6)
ix = b200ret.TermMajorIndex.from_csr(data, indices, indptr, dl, n_vocab=n_vocab, idf=idf, avgdl=avgdl)
ns, st, cap = C.c_int32(0), C.c_int32(0), C.c_int32(0)
b200ret._abi.lib.b2r_fused_plan(C.byref(ix._desc), k, C.byref(ns), C.byref(st), C.byref(cap))
a = ix.search(*q, k)
b200ret.set_approx_prefilter(False)
f = ix.search(*q, k)
torch.cuda.synchronize()
ok = bool(torch.equal(a[0], f[0]) and torch.equal(a[1], f[1]))
print("plan", ns.value, st.value, cap.value, "prefilter == f64 path:", ok)
sys.exit(0 if ok else 1)
