mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "warp_ring_every_depth" > gpurun_out/r2h_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_gputest.log
tail -15 gpurun_out/r2h_gputest.log
