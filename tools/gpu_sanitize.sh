#!/bin/bash
# compute-sanitizer over one small evaluation (smoke(): cells, neighbour pass, density, force against the oracle) with
# both neighbour kernels, and over the table-less / fallback paths of the cell-group kernel -> gpurun_out/r2_sanitizer.txt
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_sanitize.sh'
out=gpurun_out/r2_sanitizer.txt
: > $out
for tool in memcheck racecheck synccheck; do
  for tiles in 1 2; do
    echo "== compute-sanitizer --tool $tool, SPH_TILES=$tiles: __graft_entry__.smoke()" >> $out
    SPH_TILES=$tiles timeout 280 compute-sanitizer --tool $tool python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^=========     \|^$" | tail -6 >> $out
  done
done
echo "== compute-sanitizer --tool memcheck: tests/test_gpu_tiles.py -k 'group_table or capacity or cutoff'" >> $out
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_tiles.py -x -q -m gpu -k "group_table or capacity or cutoff" 2>&1 | grep -v "^=========     \|^$" | tail -6 >> $out
cat $out
