out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/f1b_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/f1b_pytest.log
bash profiles/gpu_variants.sh f1b c4 "PHI_GPU_NO_L2_PIN=1" "PHI_X=0"
bash profiles/gpu_variants.sh f1b c2 "PHI_GPU_NO_L2_PIN=1" "PHI_X=0"
