cd $GRAFT_REPO_ROOT
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r02j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02j_tests.log
tail -30 gpurun_out/r02j_tests.log
