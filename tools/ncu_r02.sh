# ncu evidence of round 2 (run under gpurun): launch list of two timed steps, full captures of the fused tower convolution and of
# the layer1 convolution.  Numbers printed by runs under ncu are never bench values.
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras --value-only"
if [ "$1" != "full-only" ]; then
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --nvtx --nvtx-include "hn_timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_step.csv $CMD > gpurun_out/ncu_list.log 2>&1
fi
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:conv_igemm_kernel<\(int\)256, \(int\)0, \(bool\)1, \(bool\)1, \(bool\)1>' -s 10 -c 2 -f -o gpurun_out/r02_conv_towers_full $CMD > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:conv_igemm_kernel<\(int\)64, \(int\)2, \(bool\)1, \(bool\)0, \(bool\)0>' -s 8 -c 2 -f -o gpurun_out/r02_conv_layer1_full $CMD > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:conv_igemm_kernel<\(int\)128, \(int\)1, \(bool\)1, \(bool\)0, \(bool\)0>' -s 6 -c 2 -f -o gpurun_out/r02_conv_layer2_full $CMD > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -n 3 gpurun_out/ncu_full1.log; tail -n 3 gpurun_out/ncu_full2.log
