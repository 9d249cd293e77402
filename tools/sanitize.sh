cd $GRAFT_REPO_ROOT
# compute-sanitizer over small parity runs (memcheck + racecheck + synccheck): evidence for profiles/
for tool in memcheck racecheck synccheck; do
  echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 3 python -m pytest tests/test_gpu_parity.py tests/test_gpu_adversarial.py -m gpu -x -q -k "kats_through_engine or quirks or typing_backwards or identical_nested" 2>&1 | grep -vE "^=========\s+(at|by|Host Frame|Device Frame|in )|^=========\s*$" | tail -12
done
