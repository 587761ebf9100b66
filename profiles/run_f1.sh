# final single-GPU cycle of round 2 (ONE gpurun command)
bash profiles/gpu_cycle2.sh f1 c4 c2
python profiles/exp_sweep.py c4 9 10 11 > gpurun_out/f1_sweep_c4.jsonl 2> gpurun_out/f1_sweep_c4.err
bash profiles/gpu_prof.sh f1 c4 "walk_sketch|read_sketch|group_count|fused_steps|chunk_key|group_fill"
bash profiles/gpu_prof.sh f1 c2 "walk_sketch|read_sketch|group_count|fused_steps|chunk_key|group_fill"
timeout 300 python profiles/readme_dropin_times.py > gpurun_out/f1_readme_dropin.json 2> gpurun_out/f1_readme_dropin.err; echo "dropin rc=$?"
