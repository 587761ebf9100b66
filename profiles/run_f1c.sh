# final single-GPU cycle of round 2 (ONE gpurun command)
bash profiles/gpu_cycle2.sh f1c c4 c2
bash profiles/gpu_prof.sh f1c c4 "fused_steps|group_count|group_fill|group_flags"
bash profiles/gpu_prof.sh f1c c2
timeout 300 python profiles/readme_dropin_times.py > gpurun_out/f1c_readme_dropin.json 2> gpurun_out/f1c_readme_dropin.err; echo "dropin rc=$?"
