#!/bin/bash
# Re-creates the evidence under profiles/ on a B200 box (run from the repo root after __graft_entry__.build()).
# Each block names the profile file(s) it produces; nothing here is needed by the tests or by bench.py.
set -e
mkdir -p gpurun_out
# r01_bench_v5.json, r01_bench_v5_reference_arm.json
python bench.py > gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json
# r01_bench_v5_n2/n4.json, r01_bench_v4_n8.json (N = 2, 4, 8)
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_n$n.json || true
done
# r01_ncu_launches_v4_power18.csv (launch list) and r01_ncu_full_v4.json / r01_ncu_full_v5_verify_kernels.json (--set full)
python bench.py --power 18 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_power18.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_power18.csv \
  python bench.py --power 18 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null
ncu --set full --clock-control none --import-source on -k regex:k_scalar_mul -s 6 -c 4 -o /tmp/ncu_smul -f \
  python bench.py --power 18 --steps 1 --warmup 3 --no-cpu-baseline --no-verify > /dev/null
ncu -i /tmp/ncu_smul.ncu-rep --page raw --csv > gpurun_out/ncu_full_scalar_mul_raw.csv
ncu --set full --clock-control none -k regex:"k_subgroup|k_msm_accumulate|k_msm_reduce" -c 14 -o /tmp/ncu_verify -f \
  python bench.py --power 18 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null
ncu -i /tmp/ncu_verify.ncu-rep --page raw --csv > gpurun_out/ncu_full_verify_raw.csv
# r01_imad_microbench_*.jsonl (multiplier microbenchmark; -DSS_MUL_INLINE for the inline row)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I snark-setup_b200/csrc -o /tmp/imad_bench tools/imad_bench.cu && /tmp/imad_bench > gpurun_out/imad_call.jsonl
# r01_extra_bench_bw6_phase2_v5.jsonl, r01_bw6_kernel_breakdown_2p16.jsonl
python tools/extra_bench.py 16 20 > gpurun_out/extra.jsonl
# r01_prepare_phase2.jsonl, r01_pairing_latency.jsonl, r01_qap_dot_product.jsonl
python tools/extra_bench.py prepare_phase2 bls12_377 20 > gpurun_out/prepare_phase2.jsonl
python tools/extra_bench.py prepare_phase2 bw6_761 16 >> gpurun_out/prepare_phase2.jsonl
python tools/extra_bench.py pairing > gpurun_out/pairing_latency.jsonl
python tools/extra_bench.py qap 20 > gpurun_out/qap.jsonl
# r01_full_config_c3_bw6_2p21.jsonl, r01_full_config_c4_bls377_2p22.jsonl, r01_host_buffer_probe.jsonl
python tools/full_configs.py c3 21 > gpurun_out/full_c3.jsonl
python tools/full_configs.py c4 22 8 > gpurun_out/full_c4.jsonl
python tools/pageable_probe.py 20 > gpurun_out/host_buffer_probe.jsonl
python tools/pageable_probe.py 22 >> gpurun_out/host_buffer_probe.jsonl
# r01_gpu_tests.log, r01_gpu_tests_full_sizes.log
python -m pytest tests -m gpu -q > gpurun_out/gpu_tests.log 2>&1
SS_TEST_FULL=1 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_full_sizes.log 2>&1
# A/B variants (profiles/r01_ab_variants.md): tools/build_variant.sh <name> "<-D flags>" then tools/ab_bench.sh

# ---- round 2 -----------------------------------------------------------------------------------------------------
# r02_bench_default_final.json, r02_bench_reference_arm_final.json, r02_bench_final_n{2,4,8}.json
python bench.py > gpurun_out/r02_bench_default_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm_final.json
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02_bench_final_n$n.json || true
done
# r02_ncu_launches_power18_final.csv (one ncu pass per box call; the same command ran plain first)
CMD="python bench.py --power 18 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
SS_CONCURRENT_VECTORS=0 $CMD > /dev/null
SS_CONCURRENT_VECTORS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
  --log-file gpurun_out/r02_ncu_launches_power18_final.csv $CMD > /dev/null
# r02_ncu_full_final.json (--set full of the hot kernels, condensed by tools/ncu_summary.py)
SS_CONCURRENT_VECTORS=0 ncu --set full --clock-control none \
  -k regex:"k_scalar_mul|k_subgroup|k_msm_accumulate|k_msm_reduce|k_decode|k_normalize|k_same_ratio" -s 40 -c 24 -o /tmp/ncu_final -f $CMD > /dev/null
ncu -i /tmp/ncu_final.ncu-rep --page raw --csv > /tmp/ncu_final_raw.csv
python tools/ncu_summary.py /tmp/ncu_final_raw.csv gpurun_out/r02_ncu_full_final.json "<command>"
# r02_ncu_full_scalar_mul_g2.json (the full-size G2 launch needs the demangled name to be selected)
SS_CONCURRENT_VECTORS=0 ncu --set full --clock-control none --kernel-name-base demangled \
  -k regex:"k_scalar_mul<ss::Bls377G2>" -c 2 -o /tmp/ncu_g2 -f $CMD > /dev/null
ncu -i /tmp/ncu_g2.ncu-rep --page raw --csv > /tmp/ncu_g2_raw.csv
python tools/ncu_summary.py /tmp/ncu_g2_raw.csv gpurun_out/r02_ncu_full_scalar_mul_g2.json "<command>"
# r02_gpu_tests_full_sizes.log
SS_TEST_FULL=1 python -m pytest tests/test_gpu_properties.py -m gpu -q > gpurun_out/r02_gpu_tests_full_sizes.log 2>&1
# r02_pairing_latency.jsonl, r02_extra_bench_bw6_phase2.jsonl
python tools/extra_bench.py pairing > gpurun_out/pairing_latency.jsonl
python tools/extra_bench.py 16 20 > gpurun_out/extra.jsonl
# strong-scaling tail: one GPU playing rank 1 of 8
python tools/shard_probe.py 22 8 1
# r02_sass_summary.md
python tools/sass_summary.py > profiles/r02_sass_summary.md
