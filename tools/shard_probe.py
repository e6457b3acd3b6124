"""One GPU playing rank `r` of `n`: device-resident timing of the verify / contribute shard of the 2^k ceremony, with the
vectors concurrent or serialised — isolates the fixed per-rank costs that limit strong scaling.
    python tools/shard_probe.py [k=22] [n=8] [r=1]"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import pyref as R
import snark_setup_b200 as S
from bench import keys

k = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
r = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
N = 1 << k
cid = S.BLS12_377
prm = S.Phase1Parameters(cid, k, 256)
acc_len, resp_len = prm.get_length(False), prm.get_length(True)
g1, g2 = R.BLS12_377.g1, R.BLS12_377.g2
g1b = torch.frombuffer(bytearray(g1.encode(g1.gen, False)), dtype=torch.uint8).to(dev)
g2b = torch.frombuffer(bytearray(g2.encode(g2.gen, False)), dtype=torch.uint8).to(dev)
blank = torch.cat([torch.zeros(64, dtype=torch.uint8, device=dev), g1b.repeat(2 * N - 1), g2b.repeat(N), g1b.repeat(N), g1b.repeat(N), g2b])
challenge = torch.empty(acc_len, dtype=torch.uint8, device=dev)
response = torch.zeros(resp_len, dtype=torch.uint8, device=dev)
newc = torch.zeros(acc_len, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream
k0, k1 = keys(b"bench-0"), keys(b"bench-1")
S.phase1_computation_dev(prm, blank.data_ptr(), acc_len, challenge.data_ptr(), acc_len, False, False, S.CHECK_NO, *k0, stream=stream)
del blank
S.phase1_computation_dev(prm, challenge.data_ptr(), acc_len, response.data_ptr(), resp_len, False, True, S.CHECK_NO, *k1, stream=stream)
torch.cuda.synchronize()
seed = hashlib.blake2b(b"bench-rho", digest_size=32).digest()


def verify(shard):
    return S.phase1_verification_vectors_dev(prm, response.data_ptr(), resp_len, True, newc.data_ptr(), acc_len, False, seed=seed,
                                             stream=stream, shard=shard, raw=True)


def contribute(shard):
    S.phase1_computation_dev(prm, challenge.data_ptr(), acc_len, response.data_ptr(), resp_len, False, True, S.CHECK_NO, *k1,
                             stream=stream, shard=shard)


def timeit(fn, reps=6):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    return round(min(ts), 2), round(sorted(ts)[len(ts) // 2], 2)


from importlib import import_module
ffi = import_module("snark-setup_b200.ffi")
out = {"k": k, "shard": [r, n]}
for conc in (True, False):
    ffi.set_concurrent_vectors(conc)
    out["verify_ms_%s" % ("concurrent" if conc else "serial")] = timeit(lambda: verify((r, n)))
    out["contribute_ms_%s" % ("concurrent" if conc else "serial")] = timeit(lambda: contribute((r, n)))
ffi.set_concurrent_vectors(True)
out["verify_whole_ms"] = timeit(lambda: verify((0, 1)), 3)
print(json.dumps(out))
