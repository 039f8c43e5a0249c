timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02_gputests_final.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_gputests_final.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "bench rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_r02_final.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['roofline']['frac'], d['clocks'], 'e2e', d['e2e']['value'], 'strong', d['strong']['ms_per_step'], 'fit', d['fit']['ms'], 'cv', d['cv']['value'], d['cv']['seconds'], 'launches', d['gpu_launches'])"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
