timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dense or tcgen05" 2>&1 | tail -15
