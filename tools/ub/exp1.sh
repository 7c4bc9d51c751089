echo "fat tiles lag1"; timeout 200 python tools/gpu_stream_probe2.py 2>&1 | tail -6
echo lag2; SDN_STREAM_LAG=2 timeout 200 python tools/gpu_stream_probe2.py 2>&1 | tail -6
echo lag3; SDN_STREAM_LAG=3 timeout 200 python tools/gpu_stream_probe2.py 2>&1 | tail -6
echo lag2 q8tn4; SDN_STREAM_LAG=2 SDN_STREAM_Q8_TN4=1 timeout 200 python tools/gpu_stream_probe2.py 2>&1 | tail -2 | head -1
