#!/bin/bash
python scratch/taper_iters_sweep.py > gpurun_out/taper_iters.log 2>&1; cat gpurun_out/taper_iters.log
AST_STFT_ITERS=8 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "stft or config1 or full_batch or config3" > gpurun_out/taper_iters_pytest.log 2>&1; tail -2 gpurun_out/taper_iters_pytest.log
