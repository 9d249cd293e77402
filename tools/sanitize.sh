cd $GRAFT_REPO_ROOT
# compute-sanitizer over a small parity run (memcheck + racecheck + synccheck): evidence for profiles/
for tool in memcheck racecheck synccheck; do
  echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kats_through_engine or quirks" 2>&1 | tail -6
done
