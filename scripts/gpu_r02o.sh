cd $GRAFT_REPO_ROOT
DCB_FORCE_WIDE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "wide or interference or datarate or golden" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('driver flags: %.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'launch ms', d['roofline']['avg_launch_ms'], d['rep_ms'], 'e2e %.3e'%d['e2e']['value'])"
timeout 300 python bench.py --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('default: %.4e'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'launch ms', d['roofline']['avg_launch_ms'], 'share', d['roofline']['kernel_share_of_step'])"
bash scripts/gpu_configs.sh r02b 2>&1 | head -4
