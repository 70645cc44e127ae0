mkdir -p gpurun_out/r3f
run() { name=$1; shift; env "$@" python profiles/step_timeline.py > gpurun_out/r3f/tl_$name.log 2>&1; echo "== $name $@"; grep "^replay" gpurun_out/r3f/tl_$name.log; grep -A5 "^kind " gpurun_out/r3f/tl_$name.log | tail -5; }
run rmw1 X=1
run rmw0 HGNN_B200_BWD_RMW=0
run c33 HGNN_B200_BWD_ENTRY_COST=0.3,0.3
run c35 HGNN_B200_BWD_ENTRY_COST=0.3,0.5
run c26 HGNN_B200_BWD_ENTRY_COST=0.2,0.6
run c515 HGNN_B200_BWD_ENTRY_COST=0.5,0.15
run c11 HGNN_B200_BWD_ENTRY_COST=1.0,1.0
run b22 HGNN_B200_BWD_BATCH=5
run b44 HGNN_B200_BWD_BATCH=3
run b28 HGNN_B200_BWD_BATCH=2
run b24 HGNN_B200_BWD_BATCH=0
