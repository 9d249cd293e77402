cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# compute-sanitizer over small parity runs of every round-2 kernel (warp-per-log incl. phase barriers and deferral, admission,
# patch stream, output packing; the c2 case runs the 8-warp team kernel): evidence for profiles/
SEL="kats_through_engine or quirks or mark_boundary_inserted_later or q4_concurrent or dense_surviving or admission_statuses or patch_kats_on_the_device or (fuzz_logs and (0 or 1)) or (generated_workloads_match_oracle and c2-24-2500)"
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_admission.py tests/test_gpu_patches.py tests/test_gpu_workloads.py -m gpu -x -q -k "$SEL" 2>&1 | grep -vE "^=========\s+(at|by|Host Frame|Device Frame|in )|^=========\s*$" | tail -14
done
