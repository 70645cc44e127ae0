mkdir -p gpurun_out/r3j
P=hgnn-2_b200
run() { name=$1; shift; env "$@" python profiles/step_timeline.py > gpurun_out/r3j/tl_$name.log 2>&1; echo "== $name $@"; grep "^replay" gpurun_out/r3j/tl_$name.log; grep -A5 "^kind " gpurun_out/r3j/tl_$name.log | tail -3 | head -2; }
run main X=1
cp $P/libhgnn_b200.so /tmp/main.so; cp $P/libhgnn_b200_alt.so $P/libhgnn_b200.so
run c3_b24 HGNN_B200_BWD_BATCH=0
run c3_auto X=1
run c3_b44 HGNN_B200_BWD_BATCH=3
run c3_b28 HGNN_B200_BWD_BATCH=2
cp /tmp/main.so $P/libhgnn_b200.so
