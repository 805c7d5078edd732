"""Same-box A/B of search-step variants on config 2 in ONE process: each variant is a build of libb200ret.so (or the same
build with another environment / bank schedule), loaded under its own file name and swapped in behind the Python mirror.
Prints one JSON line per variant: step ms (eager launches, CUDA events), dominant-kernel ms (b2r_profile_fused_ms) and a
checksum of the results (all variants must agree).

    python tools/ab_scorer.py name=path.so[,ENV=V,...][,bank=16] ...
"""
import ctypes as C
import hashlib
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np      # noqa: E402
import torch            # noqa: E402
import bench            # noqa: E402
import b200ret          # noqa: E402
from b200ret import _abi  # noqa: E402


def load(path, env):
    old = {k_: os.environ.get(k_) for k_ in env}
    os.environ.update(env)
    tmp = tempfile.NamedTemporaryFile(suffix=".so", delete=False).name
    shutil.copy(path, tmp)
    lib = C.CDLL(tmp)
    for name, (res, args) in _abi.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    for k_, v in old.items():
        if v is None:
            os.environ.pop(k_, None)
        else:
            os.environ[k_] = v
    return lib


def main():
    steps = int(os.environ.get("AB_STEPS", "20"))
    w = bench.make_workload(os.environ.get("AB_WORKLOAD", "c2"))
    k = w["k"]
    dev = torch.device("cuda", 0)
    q = [torch.from_numpy(w[n]).to(dev) for n in ("q_ptr", "q_terms", "q_w")]
    nq = len(w["q_ptr"]) - 1
    first = None
    for spec in sys.argv[1:]:
        name, rest = spec.split("=", 1)
        parts = rest.split(",")
        path, env, bank, prefilter = parts[0], {}, 1, 1
        for p in parts[1:]:
            a, b = p.split("=")
            if a == "bank":
                bank = int(b)
            elif a == "prefilter":
                prefilter = int(b)
            else:
                env[a] = b
        lib = _abi.lib = load(os.path.join(ROOT, path), env)
        lib.b2r_set_bank_schedule(bank)
        lib.b2r_set_approx_prefilter(prefilter)
        ix = b200ret.TermMajorIndex.from_csr(w["data"], w["indices"], w["indptr"], w["dl"], n_vocab=w["n_vocab"],
                                             idf=w["idf"], avgdl=w["avgdl"])
        for _ in range(5):
            idx, val = ix.search(*q, k)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            ix.search(*q, k)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        lib.b2r_set_profiling(1)
        ks = []
        for _ in range(8):
            ix.search(*q, k)
            t = C.c_float(0)
            lib.b2r_profile_fused_ms(C.byref(t), None)
            ks.append(t.value)
        lib.b2r_set_profiling(0)
        h = hashlib.sha256(idx.cpu().numpy().tobytes() + val.cpu().numpy().tobytes()).hexdigest()[:16]
        first = first or h
        print(json.dumps({"variant": name, "spec": rest, "step_ms": round(ms, 4), "kernel_ms": round(float(np.mean(ks[2:])), 4),
                          "queries_per_s": round(nq / ms * 1e3), "results": h, "same_as_first": h == first}), flush=True)
        del ix
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
