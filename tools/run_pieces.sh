#!/bin/bash
# One log with every piece timed by the tools (short bursts, not benchmarks of record): usage tools/run_pieces.sh out.log
out=${1:-gpurun_out/pieces.log}
: > "$out"
for s in "tools/time_pieces.py" "tools/time_sizes.py 1000 3000 1536 5000 96 360 2187 6000 100:262144 1025 4097 8000 8002 8191 10006 12288 13824 16000 40000 7875 12003" \
         "tools/time_next_rows.py" "tools/time_cols.py" "tools/time_clenshaw.py" "tools/time_dmma.py" "tools/time_latency.py"; do
    echo "==== python $s" >> "$out"
    python $s >> "$out" 2>&1
done
