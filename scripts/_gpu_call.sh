python scripts/host_prof.py > gpurun_out/host_prof.txt 2>&1; head -64 gpurun_out/host_prof.txt | cut -c1-170
python scripts/border_prof.py 2>&1 | head -1
