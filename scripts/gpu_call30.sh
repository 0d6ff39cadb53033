echo "== model tests with balanced partition"
PADT_MEGA_FLAGS=205 timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -x --timeout 100 2>&1 | tail -3
echo "== model tests default"
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -x --timeout 100 2>&1 | tail -3
echo "== sweep 3B"
timeout 200 python scripts/mega_sweep.py --steps 64 --reps 3 --configs "16,77,0;16,205,0;16,77,0;16,205,0" --trace "16,205,0" 2>&1 | grep -E "^\{|mega cta 0" | tail -8
