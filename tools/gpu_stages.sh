#!/bin/bash
# Bring-up runner for a GPU box: each stage in its own process (a faulting kernel poisons only its
# own CUDA context) and under its own timeout; everything is logged to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {
  name=$1; shift
  timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$*" > gpurun_out/stage_$name.log 2>&1
  echo "stage $name exit $?" | tee -a gpurun_out/stages.txt
  tail -3 gpurun_out/stage_$name.log
}
: > gpurun_out/stages.txt
run prep "prep_operand"
run exact "exact_rows"
run matrix "similarity_computer_matches or ReferenceUnitTests"
run tile "tensor_core_tile"
run bound "error_bound"
run topk2k "topk_matches_oracle_2k"
run rest "not prep_operand and not exact_rows and not similarity_computer_matches and not ReferenceUnitTests and not tensor_core_tile and not error_bound and not topk_matches_oracle_2k"
cat gpurun_out/stages.txt
