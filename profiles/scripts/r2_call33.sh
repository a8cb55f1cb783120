python profiles/trace_gemm_deep.py > gpurun_out/trace_gemm_deep.txt 2>&1; cat gpurun_out/trace_gemm_deep.txt
