python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu_final.log
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/r02_smoke_final.log
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -2 gpurun_out/r02_bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_final.json 2> gpurun_out/r02_bench_ref_final.err; head -c 300 gpurun_out/r02_bench_ref_final.json; echo
python tools/bench_conv_f16.py 2>&1 | tail -6 | tee gpurun_out/r02_bench_conv_f16.txt
for t in "tests/test_gpu_ops.py::test_dual_output_int32_and_fused_int8" "tests/test_gpu_ops.py::test_linear_over_in_memory_concat_of_features_and_occupancy_bits" "tests/test_gpu_ops.py::test_epilogue_tie_free_and_tie_chunks" "tests/test_gpu_float.py::test_tensor_core_wgrad_equals_per_offset_gemm" "tests/test_gpu_train.py::test_occupancy_net_trains_and_wgrad_paths_agree" "tests/test_gpu_ops.py::test_requant_into_column_slices_of_one_buffer"; do echo "isolated $t: $(python -m pytest "$t" -q -m gpu 2>&1 | tail -1)"; done | tee gpurun_out/r02_isolated_new_tests.txt
