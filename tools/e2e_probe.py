import sys, time, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'oracle'))
import torch, hashlib
import snark_setup_b200 as S
from snark_setup_b200 import ffi
import bench
ffi.init([0])
dev = torch.device('cuda', 0)
k = 20; N = 1 << k
prm = S.Phase1Parameters(S.BLS12_377, k, 256)
acc_len, resp_len = prm.get_length(False), prm.get_length(True)
import pyref as R
g1, g2 = R.BLS12_377.g1, R.BLS12_377.g2
g1b = torch.frombuffer(bytearray(g1.encode(g1.gen, False)), dtype=torch.uint8).to(dev)
g2b = torch.frombuffer(bytearray(g2.encode(g2.gen, False)), dtype=torch.uint8).to(dev)
blank = torch.cat([torch.zeros(64, dtype=torch.uint8, device=dev), g1b.repeat(2 * N - 1), g2b.repeat(N), g1b.repeat(N), g1b.repeat(N), g2b])
challenge = torch.empty(acc_len, dtype=torch.uint8, device=dev); response = torch.empty(resp_len, dtype=torch.uint8, device=dev)
k0, k1 = bench.keys(b"bench-0"), bench.keys(b"bench-1")
S.phase1_computation_dev(prm, blank.data_ptr(), acc_len, challenge.data_ptr(), acc_len, False, False, S.CHECK_NO, *k0)
del blank
h_in = torch.empty(acc_len, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(resp_len, dtype=torch.uint8, pin_memory=True)
h_in.copy_(challenge); torch.cuda.synchronize()
def e2e(tag):
    ffi.profile_reset(); ffi.profile_enable(True)
    torch.cuda.synchronize(); t = time.perf_counter()
    S.phase1_computation(prm, h_in.numpy(), h_out.numpy(), False, True, S.CHECK_NO, *k1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    ffi.profile_enable(False)
    p = ffi.profile_read()
    print(tag, round(dt * 1e3, 1), 'ms; kernels', {kk.replace('bls12_377.', ''): round(v['ms'], 1) for kk, v in p.items()}, flush=True)
def devstep(tag):
    torch.cuda.synchronize(); t = time.perf_counter()
    S.phase1_computation_dev(prm, challenge.data_ptr(), acc_len, response.data_ptr(), resp_len, False, True, S.CHECK_NO, *k1)
    torch.cuda.synchronize(); print(tag, round((time.perf_counter() - t) * 1e3, 1), 'ms', flush=True)
for i in range(3): e2e('e2e before verify')
for i in range(2): devstep('dev before verify')
newc = torch.empty(acc_len, dtype=torch.uint8, device=dev)
seed = bytes(32)
S.phase1_computation_dev(prm, challenge.data_ptr(), acc_len, response.data_ptr(), resp_len, False, True, S.CHECK_NO, *k1)
t = time.perf_counter()
S.phase1_verification_vectors_dev(prm, response.data_ptr(), resp_len, True, newc.data_ptr(), acc_len, False, seed=seed)
torch.cuda.synchronize(); print('verify', round((time.perf_counter() - t) * 1e3, 1), 'ms', flush=True)
for i in range(3): e2e('e2e after verify')
for i in range(2): devstep('dev after verify')
