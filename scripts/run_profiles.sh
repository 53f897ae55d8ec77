# ncu artefacts of the bench's timed region (one B200): launch list, --set full of the step kernel
# and of the tail render kernel.  Numbers printed by bench.py under ncu are not bench values.
R=${R:-r02}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# launches before the timed region: 1 post-reset + 130 burn-in x (order, step) + 3 warm-up x (order, step, render)
timeout 600 $NCU --metrics gpu__time_duration.sum -k regex:moog --launch-skip 270 --launch-count 60 --csv \
  --log-file gpurun_out/${R}_launches.csv python bench.py --no-cpu --no-clocks > gpurun_out/ncu_bench1.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:moog_step_kernel --launch-skip 136 --launch-count 1 \
  -o gpurun_out/${R}_step -f python bench.py --no-cpu --no-clocks --steps 3 > gpurun_out/ncu_bench2.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:moog_render_tail --launch-skip 4 --launch-count 1 \
  -o gpurun_out/${R}_render_tail -f python bench.py --no-cpu --no-clocks --steps 3 > gpurun_out/ncu_bench3.log 2>&1
ls -la gpurun_out/${R}_*
