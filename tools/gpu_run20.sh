cd $GRAFT_REPO_ROOT
bash tools/prof2.sh r02_team_c2 c2 1000 merge_logs_team_kernel
