tag=r3i
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-graph --no-extras --preroll 2 > gpurun_out/${tag}_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sim_dynamics|sim_minv|sim_contacts|sim_solve" -s 12 -c 4 -o gpurun_out/${tag}_sim4096_full -f python tools/prof_sim.py 4096 6 > gpurun_out/${tag}_ncu_sim4096.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sim_solve|sim_contacts" -s 790 -c 2 -o gpurun_out/${tag}_solve_contacts_full -f python tools/prof_contacts.py 350 > gpurun_out/${tag}_ncu_solve_contacts.log 2>&1
for f in sim4096_full solve_contacts_full; do ncu -i gpurun_out/${tag}_${f}.ncu-rep --page raw --csv > gpurun_out/${tag}_${f}_raw.csv 2>/dev/null; done
python tools/prof_contacts.py 900 graph 2>&1 | grep arms > gpurun_out/${tag}_contacts_over_episodes.txt
ls -la gpurun_out | grep r3i
