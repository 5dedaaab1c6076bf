mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_audio.py -q -m gpu --timeout 120 2>&1 | tail -12
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
