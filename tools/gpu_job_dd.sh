timeout 600 python -m pytest tests/test_gpu_sampler.py -x -q -m gpu -k "references_own_loop" 2>&1 | grep -E "passed|failed|^E " | head -12
