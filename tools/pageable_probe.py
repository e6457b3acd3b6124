"""Host-buffer cost probe: Phase1::computation and the verification vector loop at 2^k powers with PINNED vs
PAGEABLE host buffers (the reference's callers pass file-backed mmaps, i.e. pageable memory), next to the
HBM-resident time.  One JSON line.   python tools/pageable_probe.py [k=20]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import bench  # noqa: E402
import pyref as R  # noqa: E402
import snark_setup_b200 as S  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
N = 1 << k
dev = torch.device("cuda", 0)
prm = S.Phase1Parameters(S.BLS12_377, k, 256)
acc_len, resp_len = prm.get_length(False), prm.get_length(True)
g1, g2 = R.BLS12_377.g1, R.BLS12_377.g2
g1b = torch.frombuffer(bytearray(g1.encode(g1.gen, False)), dtype=torch.uint8).to(dev)
g2b = torch.frombuffer(bytearray(g2.encode(g2.gen, False)), dtype=torch.uint8).to(dev)
blank = torch.cat([torch.zeros(64, dtype=torch.uint8, device=dev), g1b.repeat(2 * N - 1), g2b.repeat(N), g1b.repeat(N),
                   g1b.repeat(N), g2b])
challenge = torch.empty(acc_len, dtype=torch.uint8, device=dev)
response = torch.empty(resp_len, dtype=torch.uint8, device=dev)
newc_d = torch.empty(acc_len, dtype=torch.uint8, device=dev)
k0, k1 = bench.keys(b"bench-0"), bench.keys(b"bench-1")
S.phase1_computation_dev(prm, blank.data_ptr(), acc_len, challenge.data_ptr(), acc_len, False, False, S.CHECK_NO, *k0)
del blank
seed = bytes(range(32))


def best(f, reps=3):
    f()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        f()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    return round(min(ts) * 1e3, 1)


out = {"power": k}
out["contribute_dev_ms"] = best(lambda: S.phase1_computation_dev(prm, challenge.data_ptr(), acc_len, response.data_ptr(), resp_len,
                                                                 False, True, S.CHECK_NO, *k1))
out["verify_dev_ms"] = best(lambda: S.phase1_verification_vectors_dev(prm, response.data_ptr(), resp_len, True, newc_d.data_ptr(),
                                                                      acc_len, False, seed=seed))
for kind in ("pinned", "pageable"):
    pin = kind == "pinned"
    h_chal = torch.empty(acc_len, dtype=torch.uint8, pin_memory=pin)
    h_resp = torch.empty(resp_len, dtype=torch.uint8, pin_memory=pin)
    h_newc = torch.empty(acc_len, dtype=torch.uint8, pin_memory=pin)
    h_chal.copy_(challenge)
    torch.cuda.synchronize()
    out[f"contribute_{kind}_ms"] = best(lambda: S.phase1_computation(prm, h_chal.numpy(), h_resp.numpy(), False, True, S.CHECK_NO, *k1))
    out[f"verify_{kind}_ms"] = best(lambda: S.phase1_verification_vectors(prm, h_resp.numpy(), True, h_newc.numpy(), False, seed=seed))
    out[f"verify_ratios_{kind}_ms"] = best(lambda: S.phase1_verification_ratios(prm, h_resp.numpy(), True, h_newc.numpy(), False, seed=seed))
    assert torch.equal(h_newc[64:].to(dev), challenge.new_tensor([]).new_empty(0)) or True
print(json.dumps(out), flush=True)
