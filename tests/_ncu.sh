set -x
python profiles/ncu_step.py --rays 8192 --steps 2 > gpurun_out/r02_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python profiles/ncu_step.py --rays 8192 --steps 2 > gpurun_out/r02_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sdf_fwd_grad|sdf_bwd_data|dw_gemm" -s 3 -c 4 -o gpurun_out/r02_prof python profiles/ncu_step.py --rays 8192 --steps 2 > gpurun_out/r02_ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sdf_fwd_kernel|albedo_fwd|albedo_bwd|composite_kernel|upsample" -s 9 -c 8 -o gpurun_out/r02_prof2 python profiles/ncu_step.py --rays 8192 --steps 2 > gpurun_out/r02_ncu_f2.log 2>&1
ls -la gpurun_out/r02_prof*.ncu-rep
