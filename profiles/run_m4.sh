# N=4 validation of the all-gather slice distribution (ONE gpurun --gpus 4 command)
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_multi.py -m gpu -x -q -k "multi_gpu and (synth_small or synth_dirty or mhc4 or toy)" > $out/m4_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/m4_pytest.log
bash profiles/gpu_scale.sh m4 4 c4
