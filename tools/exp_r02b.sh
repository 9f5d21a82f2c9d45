python -m pytest tests -m gpu -q --no-header -rf 2>&1 | tail -30 > gpurun_out/pytest_r02b.log; tail -3 gpurun_out/pytest_r02b.log
for prio in high same low; do for cap in 0 32 16; do
  echo "prio=$prio cap=$cap" >> gpurun_out/exp_r02b.txt
  HN_POSE_PRIO=$prio HN_POSE_CTAS=$cap python bench.py --steps 30 --warmup 5 --value-only >> gpurun_out/exp_r02b.txt 2>&1
done; done
for kb in 24 36; do
  echo "split_min_kb=$kb" >> gpurun_out/exp_r02b.txt
  HN_SPLIT_MIN_KB=$kb python bench.py --steps 30 --warmup 5 --value-only >> gpurun_out/exp_r02b.txt 2>&1
done
cat gpurun_out/exp_r02b.txt
