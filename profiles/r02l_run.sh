#!/bin/bash
# ncu evidence: launch list of the default bench line, full captures of the mesh kernels and of
# one evaluation kernel, full capture of the pipeline kernel; the reports are exported to CSV on
# the box (raw page per kernel + source-page hot spots) because only 64 MiB travel back
mkdir -p gpurun_out /tmp/ncu
T=${1:-r02l}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --single-mode > gpurun_out/${T}_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mesh_|eval_kernel' -c 12 -o /tmp/ncu/mesh python profiles/bench_mesh.py 128 > gpurun_out/${T}_mesh_ncu.log 2>&1; echo "mesh ncu rc=$?"
ncu -i /tmp/ncu/mesh.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_raw_mesh_kernels.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fit_pipeline' -s 2 -c 2 -o /tmp/ncu/pipeline python bench.py --steps 1 --warmup 1 --no-cpu-baseline --single-mode > gpurun_out/${T}_pipeline_ncu.log 2>&1; echo "pipeline ncu rc=$?"
ncu -i /tmp/ncu/pipeline.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_raw_pipeline_kernel.csv 2>/dev/null
ncu -i /tmp/ncu/pipeline.ncu-rep --page source --csv > /tmp/ncu/pipeline_source.csv 2>/dev/null
python - <<'PY'
import csv, collections, sys
# per source line: warp-stall samples and instructions executed, top 60 (second launch = main grid)
rows = list(csv.reader(open('/tmp/ncu/pipeline_source.csv', errors='ignore')))
hdr = None
agg = collections.Counter(); inst = collections.Counter()
for r in rows:
    if hdr is None:
        if 'Source' in r and any('Sampling' in c for c in r): hdr = r
        continue
    try:
        d = dict(zip(hdr, r))
        src = d.get('Source', '')
        key = [k for k in hdr if 'Warp Stall Sampling (All' in k][0]
        agg[src] += float(d[key] or 0)
        ik = [k for k in hdr if k.startswith('Instructions Executed')][0]
        inst[src] += float(d[ik] or 0)
    except Exception:
        pass
tot = sum(agg.values()) or 1
with open('gpurun_out/%s_ncu_source_hotspots_pipeline_kernel.txt' % sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/r02l_ncu_source_hotspots_pipeline_kernel.txt', 'w') as f:
    f.write('warp-stall samples per SASS-attributed source text, fit_pipeline_kernel<float> (total %d)\n' % tot)
    for s, v in agg.most_common(60):
        f.write('%8d %5.1f%% inst %10d  %s\n' % (v, 100 * v / tot, inst[s], s[:140]))
PY
ls -la gpurun_out/
