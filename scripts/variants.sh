export CUDA_LAUNCH_BLOCKING=1
run() { python scripts/try_iss.py "$@" 2>&1 | tail -1 | cut -c1-60; }
for e in 8 64 256 1024 2048 4096 8192 16384; do echo "--- extra $e"; FB_DEBUG_SMEM_EXTRA=$e run '{"words": ["[1]"], "mode": "extended", "weighting": ["Indices", {"total": true}]}' '[4,2,80]'; done
