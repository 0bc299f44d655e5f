for cfg in 6,3,32 6,2,32 5,3,32 4,3,32 4,4,32 6,4,32 8,1,32 10,0,32 6,3,16 6,3,8 6,3,64 5,4,16 7,3,32; do
  echo "cfg $cfg"; REART_KNNW_TUNE=$cfg python scripts/gpu_flow_blend_timing.py 2>&1 | grep -E "identical|windowed"
done
