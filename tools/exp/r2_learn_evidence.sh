python tools/learn_cluster_profile.py 1024 > gpurun_out/r2r_learn_timeline_warm.txt 2>&1
python tools/learn_cluster_profile.py 1024 flush > gpurun_out/r2r_learn_timeline_flushed.txt 2>&1
python tools/learn_timing.py 1024 > gpurun_out/r2r_learn_timing.txt 2>&1
RLOA_LEARN_CLUSTER=0 python tools/learn_timing.py 1024 > gpurun_out/r2r_learn_timing_multilaunch.txt 2>&1
cat gpurun_out/r2r_learn_timeline_warm.txt gpurun_out/r2r_learn_timing.txt gpurun_out/r2r_learn_timing_multilaunch.txt
