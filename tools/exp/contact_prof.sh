for v in A B; do
if [ $v = B ]; then cp tools/exp/librloa_B.so robotic_manipulator_rloa_b200/librloa_b200.so; fi
python -m pytest tests/test_sim_gpu.py -x -q -m gpu 2>&1 | tail -2
for dbg in 0 2; do
RLOA_CONTACT_DEBUG=$dbg timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sim_solve --csv --log-file gpurun_out/r2p_dur_${v}_$dbg.csv python tools/prof_contacts.py 400 > gpurun_out/r2p_log_${v}_$dbg.log 2>&1
tail -1 gpurun_out/r2p_log_${v}_$dbg.log
done; done
