for e in "B200SDR_FUSED=0" "B200SDR_FUSED=1"; do
env $e python bench.py --steps 100 --warmup 10 --warmup-seconds 0.3 --skip-e2e --skip-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$e', d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['kernel_variant'])"
done
