#!/bin/bash
# ingest -> result wall of the stand-alone CLI on configs[1]-sized files (c2: 49 walks x 240 k steps, 335,891 reads x 150 bp)
cd /root/repo
{
for kind in bgzf plain; do
  for rep in 1 2; do
    s=$(date +%s.%N)
    PHI_CLI_TIMES=1 PHI_HOST_TIMES=1 tmp_inputs/phi_index_cli -g tmp_inputs/c2.$kind.gfa.gz -r tmp_inputs/c2.$kind.fq.gz 2>&1 | grep -v " : " | sed "s/^/[$kind $rep] /"
    e=$(date +%s.%N)
    echo "[$kind $rep] process wall $(echo "$e - $s" | bc -l 2>/dev/null || python3 -c "print($e - $s)") s"
  done
done
echo "host threads: $(nproc)"
} > gpurun_out/cli_wall_c2.txt 2>&1
tail -40 gpurun_out/cli_wall_c2.txt
