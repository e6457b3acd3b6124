set -x
cd $GRAFT_REPO_ROOT
export SS_CONCURRENT_VECTORS=0
CMD="python bench.py --power 18 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_ncu_launches_power18.csv $CMD > gpurun_out/r02_ncu_l.log 2>&1
tail -2 gpurun_out/r02_ncu_l.log
$CMD > gpurun_out/r02_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_scalar_mul -s 8 -c 3 -f -o /tmp/r02_prof_smul $CMD > gpurun_out/r02_ncu_f1.log 2>&1
tail -2 gpurun_out/r02_ncu_f1.log
ncu -i /tmp/r02_prof_smul.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_smul_raw.csv
ncu -i /tmp/r02_prof_smul.ncu-rep --page source --csv > /tmp/smul_source.csv 2>/dev/null
python - <<'P'
import csv, collections, re, sys
# per-opcode executed instruction counts from the source page (one table per kernel launch)
rows = list(csv.reader(open('/tmp/smul_source.csv', errors='replace')))
hdr = None
out = []
cur = None
for r in rows:
    if len(r) > 3 and ('Source' in r and any('Instructions Executed' in c for c in r)):
        hdr = r
        cur = collections.Counter()
        out.append(cur)
        si = r.index('Source')
        ei = [i for i, c in enumerate(r) if c.strip() == 'Instructions Executed']
        ei = ei[0] if ei else None
        continue
    if hdr and cur is not None and len(r) == len(hdr) and ei is not None:
        m = re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si])
        if m:
            try:
                cur[m.group(1)] += int(float(r[ei] or 0))
            except ValueError:
                pass
with open('gpurun_out/r02_ncu_smul_opcode_mix.txt', 'w') as f:
    for k, c in enumerate(out):
        tot = sum(c.values())
        f.write(f'launch {k}: total warp instructions executed {tot}\n')
        for op, n in c.most_common(25):
            f.write(f'  {op:28s} {n:14d} {100.0*n/max(1,tot):6.2f}%\n')
print(open('gpurun_out/r02_ncu_smul_opcode_mix.txt').read()[:3000])
P
head -c 600 /tmp/smul_source.csv
$CMD > gpurun_out/r02_ncu_plain3.log 2>&1 &&
ncu --set full --clock-control none -k regex:"k_subgroup|k_msm_accumulate|k_decode|k_normalize|k_msm_reduce" -s 30 -c 16 -f -o /tmp/r02_prof_verify $CMD > gpurun_out/r02_ncu_f2.log 2>&1
tail -2 gpurun_out/r02_ncu_f2.log
ncu -i /tmp/r02_prof_verify.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_verify_raw.csv
du -sh gpurun_out
