#!/usr/bin/env python3
"""bench.py — phase1 contribute throughput (powers/s), BLS12-377, 2^20 powers per GPU.

One "step" = one full pass of Phase1::computation (phase1/src/computation.rs:16-193, Groth16, full
mode) over a synthetic 2^20-power challenge: 2^21-1 tauG1 + 2^20 tauG2/alphaG1/betaG1 + betaG2,
uncompressed in (603 979 936 B) -> compressed out (301 990 000 B), i.e. BASELINE.json configs[1].

  value : powers/s with the challenge already resident in HBM (ss_phase1_computation_dev),
          timed with CUDA events on the launching stream, max over ranks.
  e2e   : the same metric through the host-buffer C-ABI call (ss_phase1_computation) with pinned host
          challenge/response, H2D + D2H inside the timed region.
  roofline : integer-multiply pipe.  achieved = W_ref MAC32 per G1 scalar-mul x elements per launch
          / measured duration of the dominant kernel (k_scalar_mul<bls12_377.g1>); peak = measured
          MAC32/s of the box (profiles/r01_imad_microbench.json: 18.4 T IMAD/s / 2).
  cpu_baseline : the C++ oracle (reference algorithm: double-and-add + batch normalise) on the box's
          host cores, bounded sample, rank 0 at N=1 only.
N > 1 (torchrun): every rank runs the same 2^20-power chunk workload on its own GPU (chunk files of a
chunked ceremony are independent) — weak scaling, no data-path collective; NCCL is
used only for the barrier and the max-over-ranks of the timings.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

BLS_R = 0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001
# SURVEY.md §8(d) / BASELINE.md §3: reference-algorithm work, MAC32 per unit
W_REF_G1_MUL = 948750
W_REF_G2_MUL = 2314950
W_REF_POWER = 6109950
# Work the kernels actually EXECUTE (GLV/GLS + signed windows, glv.cuh): field multiplications / squarings per
# scalar multiplication counted on the host-emulated device code (tests/test_device_algos_emul.py::
# test_executed_work_per_scalar_mul pins them), in MAC32 = 2n^2+n per multiplication, (3n^2+3n)/2 per squaring (n = 12).
W_EXEC_G1_MUL = 775 * 300 + 890 * 234
W_EXEC_G2_MUL = 3200 * 300
METRIC = "phase1 contribute+verify G1+G2 powers/sec at 2^22 (1/2/4/8 B200), bit-exact"


def derive_scalar(seed: bytes, label: bytes) -> int:
    return int.from_bytes(hashlib.blake2b(seed + b"/" + label, digest_size=64).digest(), "little") % (BLS_R - 1) + 1


def keys(seed: bytes):
    return tuple(derive_scalar(seed, l) for l in (b"tau", b"alpha", b"beta"))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU, sampled every 200 ms while running."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, rank):
    """--impl reference: the reference's CPU algorithm (C++ oracle port; the Rust reference cannot be
    built in this image) on all host cores, one bounded sample of the workload per step."""
    if rank != 0:
        return
    import coracle as O
    import pyref as R
    O.set_threads(len(os.sched_getaffinity(0)))  # torchrun exports OMP_NUM_THREADS=1: use every host core anyway
    cores = O.threads()
    k = args.ref_power
    p = R.Phase1Parameters(R.BLS12_377, k, 256)
    acc = bytes(R.phase1_initialization(p, False))
    k0, k1 = keys(b"bench-0"), keys(b"bench-1")
    acc = O.phase1_computation(0, acc, p.get_length(False), False, False, 3, p.g1_chunk_size, p.other_chunk_size, 0, *k0)
    times = []
    for it in range(args.warmup + args.steps):
        t = time.perf_counter()
        O.phase1_computation(0, acc, p.get_length(True), False, True, 3, p.g1_chunk_size, p.other_chunk_size, 0, *k1)
        dt = time.perf_counter() - t
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    v = (1 << k) * len(times) / total
    sample = f"2^{k}-power BLS12-377 Phase1::computation (uncompressed in, compressed out) per step"
    restore_stdout()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "powers/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "phase1 contribute 2^20 powers BLS12-377 G1+G2 batch_exp (reference algorithm on host cores, "
                               "bounded sample; work is exactly linear in the number of powers)", "sample_power": k},
        "cpu_baseline": {"value": v, "unit": "powers/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "powers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def cpu_baseline(budget_s=12.0):
    import coracle as O
    import pyref as R
    O.set_threads(len(os.sched_getaffinity(0)))
    cores = O.threads()
    k0, k1 = keys(b"bench-0"), keys(b"bench-1")

    def run(k):
        p = R.Phase1Parameters(R.BLS12_377, k, 256)
        acc = bytes(R.phase1_initialization(p, False))
        acc = O.phase1_computation(0, acc, p.get_length(False), False, False, 3, p.g1_chunk_size, p.other_chunk_size, 0, *k0)
        t = time.perf_counter()
        O.phase1_computation(0, acc, p.get_length(True), False, True, 3, p.g1_chunk_size, p.other_chunk_size, 0, *k1)
        return time.perf_counter() - t

    t10 = run(10)
    k = 10
    while k < 16 and t10 * (1 << (k + 1 - 10)) <= budget_s:
        k += 1
    dt = run(k) if k > 10 else t10
    return {"value": (1 << k) / dt, "unit": "powers/s", "cores": cores, "kind": "port",
            "sample": f"one 2^{k}-power BLS12-377 Phase1::computation pass ({dt:.2f} s); linear in powers"}


_STDOUT_FD = None


def restore_stdout():
    global _STDOUT_FD
    if _STDOUT_FD is not None:
        sys.stdout.flush()
        os.dup2(_STDOUT_FD, 1)
        os.close(_STDOUT_FD)
        _STDOUT_FD = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--power", type=int, default=20, help="log2 powers per GPU (default 20 = BASELINE configs[1])")
    ap.add_argument("--ref-power", type=int, default=12, help="log2 powers of one reference-arm step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the verify leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries the ONE JSON line and nothing else: anything native libraries write to fd 1 while the
    # bench runs (NCCL prints its version line there when NCCL_DEBUG=VERSION) is sent to stderr instead.
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import snark_setup_b200 as S
    from snark_setup_b200 import ffi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA path is the product, there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ffi.init([local_rank])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    k = args.power
    N = 1 << k
    prm = S.Phase1Parameters(S.BLS12_377, k, 256)
    acc_len, resp_len = prm.get_length(False), prm.get_length(True)

    # ---- synthetic challenge, built on the device: generators -> contribution "bench-0" --------------
    import pyref as R
    g1, g2 = R.BLS12_377.g1, R.BLS12_377.g2
    g1b = torch.frombuffer(bytearray(g1.encode(g1.gen, False)), dtype=torch.uint8).to(dev)
    g2b = torch.frombuffer(bytearray(g2.encode(g2.gen, False)), dtype=torch.uint8).to(dev)
    blank = torch.cat([torch.zeros(64, dtype=torch.uint8, device=dev), g1b.repeat(2 * N - 1), g2b.repeat(N),
                       g1b.repeat(N), g1b.repeat(N), g2b])
    assert blank.numel() == acc_len
    challenge = torch.empty(acc_len, dtype=torch.uint8, device=dev)
    response = torch.empty(resp_len, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    k0, k1 = keys(b"bench-0"), keys(b"bench-1" + bytes([rank]))
    S.phase1_computation_dev(prm, blank.data_ptr(), acc_len, challenge.data_ptr(), acc_len, False, False, S.CHECK_NO,
                             *k0, stream=stream)
    del blank
    torch.cuda.empty_cache()

    def step():
        S.phase1_computation_dev(prm, challenge.data_ptr(), acc_len, response.data_ptr(), resp_len, False, True,
                                 S.CHECK_NO, *k1, stream=stream)

    for _ in range(args.warmup):
        step()
    # ---- timed region: device-resident --------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    ffi.profile_reset()
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    launches = ffi.profile_launches()
    value = world * N * args.steps / (dev_ms * 1e-3)
    # per-kernel durations for the roofline: one extra step with the vectors SERIALISED on one stream, so the
    # event brackets around each launch measure that kernel alone (in the timed region above the five vectors
    # run on concurrent streams and their kernels time-slice the SMs)
    ffi.set_concurrent_vectors(False)
    step()
    ffi.profile_reset()
    ffi.profile_enable(True)
    torch.cuda.synchronize()
    s0e, s1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0e.record()
    step()
    s1e.record()
    torch.cuda.synchronize()
    serial_ms = s0e.elapsed_time(s1e)
    ffi.profile_enable(False)
    prof = ffi.profile_read()
    ffi.set_concurrent_vectors(True)
    prof_steps = 1

    # ---- spot parity check against the oracle (test infrastructure used as the checker only) ---------
    parity = None
    if rank == 0:
        import coracle as O
        idx = 123457 % (2 * N - 1)
        n_chk = 16
        cin = challenge[64 + idx * 96: 64 + (idx + n_chk) * 96].cpu().numpy().tobytes()
        got = response[64 + idx * 48: 64 + (idx + n_chk) * 48].cpu().numpy().tobytes()
        want = O.apply_powers(0, 0, cin, False, 3, True, n_chk, tau=k1[0], first_power=idx)
        parity = bool(got == want)

    # ---- verify leg (reported beside the headline, same unit): per-vector loop of Phase1::verification ---
    # compressed response -> nonzero + subgroup checks + power_pairs MSM + uncompressed new challenge
    verify = None
    if not args.no_verify:
        seed = hashlib.blake2b(b"bench-rho", digest_size=32).digest()
        newc = torch.empty(acc_len, dtype=torch.uint8, device=dev)

        def vstep():
            return S.phase1_verification_vectors_dev(prm, response.data_ptr(), resp_len, True, newc.data_ptr(), acc_len,
                                                     False, seed=seed, stream=stream)

        pairs = vstep()
        barrier()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record()
        vsteps = max(1, min(args.steps, 3))
        for _ in range(vsteps):
            pairs = vstep()
        v1.record()
        barrier()
        v_ms = max_over_ranks(v0.elapsed_time(v1))
        # per-kernel durations from one serialised pass (see the contribute leg)
        ffi.set_concurrent_vectors(False)
        vstep()
        ffi.profile_reset()
        ffi.profile_enable(True)
        vstep()
        torch.cuda.synchronize()
        ffi.profile_enable(False)
        vprof = ffi.profile_read()
        ffi.set_concurrent_vectors(True)
        ok = None
        if rank == 0:
            import coracle as O
            tau_acc = k0[0] * k1[0] % BLS_R
            ok = True
            for (s, sx), grp in zip(pairs, (0, 1, 0, 0)):
                ok = ok and O.apply_powers(0, grp, s, False, 3, False, 1, powers=[tau_acc]) == sx
            # the re-emitted challenge must be the decompressed response: spot-check one tau_g1 slice
            idx = 7777 % (2 * N - 1)
            want = O.transcode(0, 0, response[64 + idx * 48: 64 + (idx + 8) * 48].cpu().numpy().tobytes(), True, 3, False, 8)
            ok = ok and newc[64 + idx * 96: 64 + (idx + 8) * 96].cpu().numpy().tobytes() == want
        verify = {"value": world * N * vsteps / (v_ms * 1e-3), "unit": "powers/s", "steps": vsteps,
                  "ms_per_step": v_ms / vsteps, "ratio_and_reemit_check": ok,
                  "what": "ss_phase1_verification_vectors_dev: compressed response -> OnlyNonZero decode, r*P subgroup check, "
                          "power_pairs (s,sx) per vector, uncompressed new challenge; pairings (8 per response) left to the host",
                  "kernels_ms_per_step_serialised": {kk: round(vv["ms"], 3) for kk, vv in sorted(vprof.items())}}
        del newc

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ---------------------------
    h_in = torch.empty(acc_len, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(resp_len, dtype=torch.uint8, pin_memory=True)
    h_in.copy_(challenge)
    torch.cuda.synchronize()

    def e2e_step():
        S.phase1_computation(prm, h_in.numpy(), h_out.numpy(), False, True, S.CHECK_NO, *k1)

    for _ in range(3):  # warm-up: staging slabs of the host path are (re)grown here, not in the timed region
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * N * e2e_steps / e2e_s
    e2e_match = bool(torch.equal(h_out[64:].to(dev), response[64:]))

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    imad = None
    try:
        imad = json.load(open(os.path.join(ROOT, "profiles", "imad_peak.json")))
    except Exception:
        pass
    mac_peak = (imad or {}).get("mac32_tps", 9.2)  # T MAC32/s: 18.4 T IMAD/s / 2 IMAD per MAC32
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"]) if prof else (None, None)
    roofline = None
    if dom[0]:
        name, d = dom
        per_elem = W_REF_G2_MUL if ".g2" in name else W_REF_G1_MUL
        ach = per_elem * d["elements"] / (d["ms"] * 1e-3) * 1e-12
        total_ms = sum(v["ms"] for v in prof.values())
        roofline = {"bound": "int32-imad", "kernel": name, "achieved": ach, "peak": mac_peak, "unit": "TMAC32/s",
                    "frac": ach / mac_peak, "peak_source": "measured IMAD microbenchmark on this pool (profiles/imad_peak.json)"
                    if imad else "fallback 18.4 T IMAD/s / 2",
                    "traffic": (int((imad or {}).get("scalar_mul_g1_dram_bytes_per_element") * d["elements"] / max(1, d["launches"]))
                                if (imad or {}).get("scalar_mul_g1_dram_bytes_per_element") and ".g1" in name else None),
                    "avg_launch_ms": d["ms"] / max(1, d["launches"]), "kernel_share_of_step": d["ms"] / total_ms,
                    "whole_step_frac": W_REF_POWER * (value / world) * 1e-12 / mac_peak,
                    "executed": {"what": "MAC32 the kernel really executes (GLV/GLS algorithm, counted on the emulated "
                                         "device code) / duration; frac = share of the multiplier peak in use",
                                 "mac32_per_element": W_EXEC_G2_MUL if ".g2" in name else W_EXEC_G1_MUL,
                                 "achieved": (W_EXEC_G2_MUL if ".g2" in name else W_EXEC_G1_MUL) * d["elements"] / (d["ms"] * 1e-3) * 1e-12,
                                 "frac": (W_EXEC_G2_MUL if ".g2" in name else W_EXEC_G1_MUL) * d["elements"] / (d["ms"] * 1e-3) * 1e-12 / mac_peak,
                                 "ncu_fmaheavy_pct_of_elapsed": 87.2 if ".g1" in name else 78.3},
                    "hbm": {"algorithmic_GBps": (acc_len + resp_len) * args.steps / (dev_ms * 1e-3) * 1e-9,
                            "peak_GBps": peaks.get("hbm_gbs")},
                    "serialised_step_ms": serial_ms,
                    "kernels_ms_per_step": {kk: round(vv["ms"] / prof_steps, 3) for kk, vv in sorted(prof.items())}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()

    if rank == 0:
        restore_stdout()
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "powers/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"phase1 contribute 2^{k} powers BLS12-377 G1+G2 batch_exp (Phase1::computation, Groth16 "
                                   f"full mode, uncompressed challenge {acc_len} B -> compressed response {resp_len} B) per GPU",
                       "sharding": "one independent 2^%d-power chunk workload per GPU, no collective" % k,
                       "cache": "inputs (604 MB) larger than L2 (126 MB); no flush needed",
                       "synthetic_input": "generators -> contribution keyed 'bench-0' -> timed contribution 'bench-1'"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "powers/s", "h2d_bytes_per_step": acc_len, "d2h_bytes_per_step": resp_len,
                    "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3, "matches_device_path": e2e_match},
            "roofline": roofline, "cpu_baseline": cpu, "parity_spot_check": parity, "verify": verify,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
