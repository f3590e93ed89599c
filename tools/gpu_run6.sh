cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02i_tests.log
tail -12 gpurun_out/r02i_tests.log
for f in small big; do for n in 16384 65536; do REXSIM_FORCE_BUILD=$f python tools/prof_cfg.py C4 $n 20 | sed "s/^/$f /"; done; done
for f in small big; do REXSIM_FORCE_BUILD=$f python tools/prof_cfg.py C2 8192 20 | sed "s/^/$f /"; done
python tools/prof_cfg.py C5 16384 20
