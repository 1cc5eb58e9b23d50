mkdir -p gpurun_out
( DEPTHPRO_ATTN_EXP=13 timeout 120 scripts/ubench/attn_prof ) > gpurun_out/attn_trace_v13.log 2>&1; echo "exit $?"
cat gpurun_out/attn_trace_v13.log
