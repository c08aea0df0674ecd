#!/bin/bash
# compute-sanitizer (memcheck) over the C boundary program and a slice of the GPU suite that exercises every kernel family at small sizes.
mkdir -p gpurun_out
gcc -std=c99 -O1 tests/cabi/boundary.c -I include -I /usr/local/cuda/include -L lattigo_b200/lib -llattigo_b200 -L /usr/local/cuda/lib64 -lcudart -lpthread \
    -Wl,-rpath,$PWD/lattigo_b200/lib -Wl,-rpath,/usr/local/cuda/lib64 -o /tmp/boundary || exit 1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 /tmp/boundary > gpurun_out/sanitizer_boundary.log 2>&1; echo "boundary rc=$?" | tee -a gpurun_out/sanitizer_boundary.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --target-processes all python -m pytest -x -q -m gpu \
    tests/test_gpu_lintrans.py::test_lintrans_lower_matrix_level_and_inplace tests/test_gpu_rgsw.py::test_external_product_fused_sizes \
    "tests/test_gpu_encryptor.py::test_gen_evaluation_key_matches_the_oracle" tests/test_gpu_keyswitch.py::test_gadget_product_fused_pipeline_logn13 \
    tests/test_gpu_ring.py::test_golden_vectors_through_gpu tests/test_gpu_wire.py > gpurun_out/sanitizer_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/sanitizer_pytest.log
tail -5 gpurun_out/sanitizer_boundary.log; grep -E "ERROR SUMMARY|passed|failed|error" gpurun_out/sanitizer_pytest.log | tail -8
