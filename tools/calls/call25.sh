timeout 100 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "pair_ or late_rescale or full_window or kv_chunks8" 2>&1 | tail -6 | cut -c1-300
